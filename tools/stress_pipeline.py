"""Race / noise check of the pipelined host API (ian_reconstruct_submit/_wait).

Small batches go through the atomic split-K path, whose fp32 summation order varies run to run; this prints
the run-to-run spread of the synchronous call next to the pipelined-vs-synchronous difference, so a real race
(differences far above the spread, or any difference at a batch size that uses no atomics) stands out.
"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import weights as ow
pkg = importlib.import_module("neural-photo-editor_b200")
m = pkg.IAN("IAN_simple.py", True, weights=ow.make_simple_weights(0))
rng = np.random.default_rng(9)
for n, reps in ((5, 40), (64, 20), (256, 20)):
    batches = [rng.uniform(-1, 1, (n, 3, 64, 64)).astype(np.float32) for _ in range(6)]
    ref = [m.reconstruct(b) for b in batches]
    sync_spread = 0.0
    for _ in range(reps):
        for b, r in zip(batches, ref):
            sync_spread = max(sync_spread, float(np.abs(m.reconstruct(b) - r).max()))
    pipe = 0.0
    for _ in range(reps):
        for got, r in zip(m.reconstruct_stream(iter(batches)), ref):
            pipe = max(pipe, float(np.abs(got - r).max()))
    zs, zp = 0.0, 0.0
    zref = [m.encode_images(b) for b in batches]
    out = m.pinned_empty((n, 3, 64, 64)); zo = m.pinned_empty((n, 100))
    for _ in range(reps):
        for b, r in zip(batches, zref):
            zs = max(zs, float(np.abs(m.encode_images(b) - r).max()))
            m.reconstruct_wait(m.reconstruct_submit(b, out, zo))
            zp = max(zp, float(np.abs(zo - r).max()))
    print("n=%3d  x_hat: sync-vs-sync %.3g  pipelined-vs-sync %.3g   z: sync %.3g  pipelined %.3g"
          % (n, sync_spread, pipe, zs, zp), flush=True)

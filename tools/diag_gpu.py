"""GPU diagnostic (not a test): error statistics of the CUDA path against the float64 oracle."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ian_numpy as on  # noqa: E402
from oracle import weights as ow  # noqa: E402

npe = importlib.import_module("neural-photo-editor_b200")
P = ow.make_simple_weights(0)
m = npe.IAN("IAN_simple.py", True, weights=P)
rng = np.random.default_rng(42)
N = int(os.environ.get("DIAG_N", "48"))
x = rng.uniform(-1, 1, (N, 3, 64, 64)).astype(np.float32)
z0 = rng.standard_normal((N, 100)).astype(np.float32)
side = rng.integers(1, 18, N)
c1 = np.array([rng.integers(0, 64 - s + 1) for s in side]); r1 = np.array([rng.integers(0, 64 - s + 1) for s in side])
boxes = np.stack([c1, r1, c1 + side, r1 + side], 1).astype(np.int32)
rgb = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
t = time.time()
zr = on.simple_encode(P, x); xr = on.simple_decode(P, z0); gr = on.simple_grad_batched(P, z0, boxes, rgb)
print("oracle %.1fs" % (time.time() - t))
for path in ("simt", "tc"):
    m.set_path(path)
    z = m.encode_images(x); xh = m.sample_at(z0); g = m.grad(z0, boxes, rgb)
    ez = np.abs(z - zr).max(axis=1); ex = np.abs(xh - xr).reshape(N, -1).max(axis=1)
    eg = np.abs(g - gr).max(axis=1) / np.abs(gr).max(axis=1)
    print(path, "z err max %.3g med %.3g | x err max %.3g med %.3g" % (ez.max(), np.median(ez), ex.max(), np.median(ex)))
    print(path, "grad rel err: med %.3g  p90 %.3g  max %.3g ; sorted top5 %s" % (np.median(eg), np.quantile(eg, .9), eg.max(), np.sort(eg)[-5:]))

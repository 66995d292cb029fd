"""diagnostic: IANv1 brush gradient, simt vs tc vs float64 autograd, by box position and batch composition"""
import importlib, sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import weights as ow, ian_torch as ot
pkg = importlib.import_module("neural-photo-editor_b200")
P = ow.make_v1_weights(0)
P64 = ot.to_torch(P, torch.float64)
rng = np.random.default_rng(8)
zb = rng.standard_normal((3, 100)).astype(np.float32)
boxes = {"interior_big": [12, 14, 44, 40], "interior_small": [30, 30, 34, 35], "border_tl": [0, 0, 9, 9], "border_br": [50, 50, 64, 64],
         "case0": [3, 5, 20, 17], "case2": [0, 47, 64, 64]}
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for path in ("tc", "simt"):
    m = pkg.IAN("IANv1.py", True, weights=P, path=path)
    for name, b in boxes.items():
        for k in (0, 2):
            zt = torch.from_numpy(zb[k:k + 1].astype(np.float64))
            ref = ot.imgrad(P64, b[0], b[1], b[2], b[3], zt, decode_fn=ot.v1_decode).numpy()
            g1 = m.grad(zb[k:k + 1], np.array([b], np.int32), None)                       # alone (batch 1)
            g3 = m.grad(zb, np.array([b, b, b], np.int32), None)[k:k + 1]                 # inside a batch of 3
            print(path, name, "sample", k, "alone %.2e" % rel(g1, ref), "in-batch %.2e" % rel(g3, ref), flush=True)
    m.close()

"""CPU ORACLE helper (test infrastructure): seeded SYNTHETIC weights in the reference checkpoint format.

The trained blobs IAN_simple.npz / IANv1.npz are git-LFS pointers in the reference, so all
parity runs use synthetic weights.  Parameter names/shapes follow SURVEY Appendix A (Lasagne
"<layer name>.<param>", IAN_simple.py:73-181) and the on-disk format follows
GANcheckpoints.save_weights (GANcheckpoints.py:11-30): np.savez_compressed of {name: ndarray}
plus a pickled 'metadata' entry.

Distributions: conv/deconv/dense W ~ N(0, 0.02) (initmethod(0.02),
IAN_simple.py:79); enc_conv1.b ~ N(0, 0.02); BN gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1),
inv_std~U(0.5,2)  -- deliberately NOT the Lasagne defaults (0/1), which would hide BN bugs.
"""
from __future__ import annotations

import pickle

import numpy as np

SIMPLE_SHAPES = [
    ("enc_conv1.W", (128, 3, 5, 5)), ("enc_conv1.b", (128,)),
    ("enc_conv2.W", (256, 128, 5, 5)), ("bnorm2", 256),
    ("enc_conv3.W", (512, 256, 5, 5)), ("bnorm3", 512),
    ("enc_conv4.W", (1024, 512, 5, 5)), ("bnorm4", 1024),
    ("enc_fc1.W", (16384, 1000)), ("bnorm_enc_fc1", 1000),
    ("enc_mu.W", (1000, 100)), ("mu_bnorm", 100),
    ("enc_logsigma.W", (1000, 100)), ("ls_bnorm", 100),
    ("l_dec_fc2.W", (100, 16384)), ("bnorm_dec_fc2", 16384),
    ("dec_conv1.W", (1024, 512, 5, 5)), ("bnorm_dc1", 512),
    ("dec_conv2.W", (512, 256, 5, 5)), ("bnorm_dc2", 256),
    ("dec_conv3.W", (256, 128, 5, 5)), ("bnorm_dc3", 128),
    ("dec_out.W", (128, 3, 5, 5)),
]


def make_simple_weights(seed=0, w_std=0.02):
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp in SIMPLE_SHAPES:
        if isinstance(shp, int):
            P[name + ".gamma"] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
            P[name + ".beta"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".mean"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".inv_std"] = rng.uniform(0.5, 2.0, shp).astype(np.float32)
        else:
            P[name] = rng.normal(0, w_std, shp).astype(np.float32)
    return P


def save_checkpoint(fname, P, metadata=None):
    """GANcheckpoints.save_weights format (GANcheckpoints.py:11-30)."""
    d = dict(P)
    d["metadata"] = np.frombuffer(pickle.dumps(metadata or {"epoch": 0, "itr": 0}, protocol=2),
                                  dtype=np.uint8)
    np.savez_compressed(fname, **d)


# ---- full IAN (reference IAN.py:67-228) -------------------------------------------------------------
def _mdcl_shapes(name, F, C, scales):
    out = [(name + "W", (F, C, 3, 3)), (name + "_coeff_base", ("coeff", F))]
    for s in scales:
        out.append((name + ("_coeff_1x1" if s == 0 else "_coeff_%d" % s), ("coeff", F)))
    return out


def _mdblock_shapes(name, F, scales):
    return (_mdcl_shapes(name, F, F, scales) + _mdcl_shapes(name + "2", F, F, scales) +
            [(name + "bnorm0", F), (name + "bnorm1", F), (name + "bnorm2", F)])


def full_shapes():
    enc = [s for s in SIMPLE_SHAPES if s[0].startswith(("enc_", "bnorm2", "bnorm3", "bnorm4", "bnorm_enc", "mu_", "ls_"))]
    made = []
    for name in ("l_IAF_mu", "l_IAF_ls"):
        for sub in ("_input", "_output_W", "_output_D"):
            made += [(name + sub + ".W", ("made", (100, 100))), (name + sub + ".b", ("made", (100,)))]
    dec = [("l_dec_fc2.W", (100, 8192)), ("l_dec_fc2.b", (8192,)), ("dec_conv1.W", (512, 512, 5, 5))]
    dec += _mdblock_shapes("dec_conv2a", 512, [0, 2]) + [("dec_conv2.W", (512, 256, 5, 5))]
    dec += _mdblock_shapes("dec_conv3a", 256, [0, 2, 3]) + [("dec_conv3.W", (256, 128, 5, 5))]
    dec += _mdblock_shapes("dec_conv4a", 128, [0, 2, 3]) + [("dec_conv4.W", (128, 128, 5, 5)), ("bnorm_dc4", 128)]
    sc = [2, 3, 4]
    dec += (_mdcl_shapes("R", 2, 128, sc) + _mdcl_shapes("G_a", 2, 128, sc) + _mdcl_shapes("G_b", 2, 2, sc) +
            _mdcl_shapes("B_a", 2, 128, sc) + _mdcl_shapes("B_b", 2, 4, sc))
    return enc + made + dec


def make_full_weights(seed=0, w_std=0.02):
    """Synthetic weights for the IAN.py graph.  MDCL coefficients ~U(0.1,0.5) (reference init is the constant
    1/(1+len(scales)), layers.py:214: randomised so that a swapped coefficient shows); MADE W,b ~N(0,0.1)."""
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp in full_shapes():
        if isinstance(shp, int):
            P[name + ".gamma"] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
            P[name + ".beta"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".mean"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".inv_std"] = rng.uniform(0.5, 2.0, shp).astype(np.float32)
        elif shp[0] == "coeff":
            P[name] = rng.uniform(0.1, 0.5, shp[1]).astype(np.float32)
        elif shp[0] == "made":
            P[name] = rng.normal(0, 0.1, shp[1]).astype(np.float32)
        elif name in ("enc_conv1.b", "l_dec_fc2.b"):
            P[name] = rng.normal(0, 0.02, shp).astype(np.float32)
        else:
            # 3x3 MDC filters get a larger std so that the residual branch is not negligible next to x
            std = 0.05 if (name.endswith("W") and not name.endswith(".W")) else w_std
            P[name] = rng.normal(0, std, shp).astype(np.float32)
    return P


# ---- IANv1 (reference IANv1.py:63-222) ----------------------------------------------------------------
def v1_shapes():
    full = full_shapes()
    enc_made = [s for s in full if s[0].startswith(("enc_", "bnorm2", "bnorm3", "bnorm4", "bnorm_enc", "mu_", "ls_", "l_IAF"))]
    dec = [("l_dec_fc2.W", (100, 16384)), ("l_dec_fc2.b", (16384,)),
           ("dec_conv1.W", (1024, 512, 5, 5)), ("bnorm_dc1", 512), ("dec_conv2.W", (512, 256, 5, 5)), ("bnorm_dc2", 256),
           ("dec_conv3.W", (256, 128, 5, 5)), ("bnorm_dc3", 128), ("dec_conv4.W", (128, 64, 5, 5)), ("bnorm_dc4", 64)]
    sc = [2, 3, 4]
    dec += (_mdcl_shapes("R", 2, 64, sc) + _mdcl_shapes("G_a", 2, 64, sc) + _mdcl_shapes("G_b", 2, 2, sc) +
            _mdcl_shapes("B_a", 2, 64, sc) + _mdcl_shapes("B_b", 2, 4, sc))
    return enc_made + dec


def make_v1_weights(seed=0, w_std=0.02):
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp in v1_shapes():
        if isinstance(shp, int):
            P[name + ".gamma"] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
            P[name + ".beta"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".mean"] = rng.normal(0, 0.1, shp).astype(np.float32)
            P[name + ".inv_std"] = rng.uniform(0.5, 2.0, shp).astype(np.float32)
        elif shp[0] == "coeff":
            P[name] = rng.uniform(0.1, 0.5, shp[1]).astype(np.float32)
        elif shp[0] == "made":
            P[name] = rng.normal(0, 0.1, shp[1]).astype(np.float32)
        elif name in ("enc_conv1.b", "l_dec_fc2.b"):
            P[name] = rng.normal(0, 0.02 if name == "enc_conv1.b" else 0.2, shp).astype(np.float32)
        else:
            std = 0.05 if (name.endswith("W") and not name.endswith(".W")) else (0.1 if name == "l_dec_fc2.W" else w_std)
            P[name] = rng.normal(0, std, shp).astype(np.float32)
    return P


# ---- BASELINE config 4 inputs: the latent-brush edit loop -----------------------------------
def config4_inputs(n=128):
    """z ~ N(0,1) seed 2; per-sample brush boxes by NPE's law (NPE.py:143-156: side w in 1..17, origin uniform so that
    the box stays inside the 64x64 frame) and target colours U(-1,1)^3, seed 3.  Returns z (n,100) f32,
    boxes (n,4) int32 [c1,r1,c2,r2], rgb (n,3) f32."""
    z = np.random.default_rng(2).standard_normal((n, 100)).astype(np.float32)
    r3 = np.random.default_rng(3)
    side = r3.integers(1, 18, n)
    c1 = np.array([r3.integers(0, 64 - s + 1) for s in side])
    r1 = np.array([r3.integers(0, 64 - s + 1) for s in side])
    boxes = np.stack([c1, r1, c1 + side, r1 + side], 1).astype(np.int32)
    rgb = r3.uniform(-1, 1, (n, 3)).astype(np.float32)
    return z, boxes, rgb

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

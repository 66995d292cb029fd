"""CPU ORACLE (test infrastructure, NOT product code) -- independent float32 torch-CPU restatement.

PARITY UNPINNED (see oracle/ian_numpy.py header).  Written against torch.nn.functional so that it
shares no arithmetic code with the float64 definitional version; the two must agree before either
is trusted (tests/test_oracle.py).  Also the timed CPU baseline of bench.py ("CPU restatement of the
reference graph, N cores" -- never "Theano").

Reference lines followed: IAN_simple.py:56-241 (graph), layers.py:419-483 (sample / deconv),
API.py:40-64 (functions and T.grad), NPE.py:199-209 (step rule).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_torch(P, dtype=torch.float32):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in P.items()
            if isinstance(v, np.ndarray) and v.dtype.kind == "f"}


def _bn(P, name, x):
    shp = (1, -1) + (1,) * (x.dim() - 2)
    s = (P[name + ".gamma"] * P[name + ".inv_std"]).reshape(shp)
    return (x - P[name + ".mean"].reshape(shp)) * s + P[name + ".beta"].reshape(shp)


def _lrelu(x):                     # lasagne LeakyRectify(0.2): 0.6x + 0.4|x|  (C.5)
    return 0.6 * x + 0.4 * x.abs()


def _relu(x):                      # lasagne rectify: 0.5(x+|x|)
    return 0.5 * (x + x.abs())


def deconv(x, W):
    """DeconvLayer (layers.py:436-483), cuDNN formulation (IAN_simple.py:141-181); C.4."""
    return F.conv_transpose2d(x, W.flip(2, 3), stride=2, padding=2, output_padding=1)


def deconv_tc2d_slice(x, W):
    """non-cuDNN formulation: TransposedConv2DLayer(crop=1, flip_filters=False) + SliceLayer[1:,1:]
    (IAN_simple.py:183-223).  KAT: must equal deconv()."""
    y = F.conv_transpose2d(x, W.flip(2, 3), stride=2, padding=1)
    return y[:, :, 1:, 1:]


def encode_mu_ls(P, x):
    h = _lrelu(F.conv2d(x, P["enc_conv1.W"], P["enc_conv1.b"], stride=2, padding=2))
    h = _lrelu(_bn(P, "bnorm2", F.conv2d(h, P["enc_conv2.W"], None, stride=2, padding=2)))
    h = _lrelu(_bn(P, "bnorm3", F.conv2d(h, P["enc_conv3.W"], None, stride=2, padding=2)))
    h = _lrelu(_bn(P, "bnorm4", F.conv2d(h, P["enc_conv4.W"], None, stride=2, padding=2)))
    h = F.elu(_bn(P, "bnorm_enc_fc1", h.flatten(1) @ P["enc_fc1.W"]))
    mu = _bn(P, "mu_bnorm", h @ P["enc_mu.W"])
    ls = _bn(P, "ls_bnorm", h @ P["enc_logsigma.W"])
    return mu, ls


def encode(P, x, deterministic=True, eps=None):
    mu, ls = encode_mu_ls(P, x)
    if deterministic:
        return mu
    return mu + torch.exp(ls) * eps


def decode(P, z, deconv_fn=deconv):
    h = _relu(_bn(P, "bnorm_dec_fc2", z @ P["l_dec_fc2.W"])).reshape(-1, 1024, 4, 4)
    h = _relu(_bn(P, "bnorm_dc1", deconv_fn(h, P["dec_conv1.W"])))
    h = _relu(_bn(P, "bnorm_dc2", deconv_fn(h, P["dec_conv2.W"])))
    h = _relu(_bn(P, "bnorm_dc3", deconv_fn(h, P["dec_conv3.W"])))
    return torch.tanh(deconv_fn(h, P["dec_out.W"]))


def imgrad(P, c1, r1, c2, r2, z):
    """API.py:59 -- T.grad(T.mean(X_hat[0,:,r1:r2,c1:c2]), Z) via autograd."""
    z = z.clone().requires_grad_(True)
    loss = decode(P, z)[0, :, int(r1):int(r2), int(c1):int(c2)].mean()
    (g,) = torch.autograd.grad(loss, z)
    return g


def imgradRGB(P, c1, r1, c2, r2, RGB, z):
    """API.py:64 -- T.grad(T.mean(sqr(-X_hat[0,:,box] + RGB[0,:,box])), Z) via autograd."""
    z = z.clone().requires_grad_(True)
    r1, r2, c1, c2 = int(r1), int(r2), int(c1), int(c2)
    loss = ((-decode(P, z)[0, :, r1:r2, c1:c2] + RGB[0, :, r1:r2, c1:c2]) ** 2).mean()
    (g,) = torch.autograd.grad(loss, z)
    return g


def grad_batched(P, z, boxes, rgb=None):
    """per-sample boxes; rgb (N,3) colour or None for the lighten gradient (see ian_numpy)."""
    z = z.clone().requires_grad_(True)
    xh = decode(P, z)
    loss = 0.0
    for k in range(z.shape[0]):
        c1, r1, c2, r2 = [int(v) for v in boxes[k]]
        patch = xh[k, :, r1:r2, c1:c2]
        if rgb is None:
            loss = loss + patch.mean()
        else:
            loss = loss + ((rgb[k].reshape(3, 1, 1) - patch) ** 2).mean()
    (g,) = torch.autograd.grad(loss, z)
    return g


def edit_loop(P, z, boxes, rgb, n_steps=32, weight=0.05):
    """NPE.py:199-209 per sample, float32 state (BASELINE config 4)."""
    z = z.clone()
    fac = (1.0 + (boxes[:, 2] - boxes[:, 0]).to(z.dtype))[:, None]
    for _ in range(n_steps):
        g = grad_batched(P, z, boxes, rgb)
        z = z - weight * g * fac
    return z

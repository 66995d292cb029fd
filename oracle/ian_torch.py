"""CPU ORACLE (test infrastructure, NOT product code) -- independent float32 torch-CPU restatement.

PARITY: pinned through oracle/ian_numpy.py, which matches the executed reference (see its header).  Written against torch.nn.functional so that it
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


def imgrad(P, c1, r1, c2, r2, z, decode_fn=None):
    """API.py:59 -- T.grad(T.mean(X_hat[0,:,r1:r2,c1:c2]), Z) via autograd.  `decode_fn` selects the graph
    (default: IAN_simple's decoder; full_decode / v1_decode for the IAN.py / IANv1.py graphs)."""
    z = z.clone().requires_grad_(True)
    loss = (decode_fn or decode)(P, z)[0, :, int(r1):int(r2), int(c1):int(c2)].mean()
    (g,) = torch.autograd.grad(loss, z)
    return g


def imgradRGB(P, c1, r1, c2, r2, RGB, z, decode_fn=None):
    """API.py:64 -- T.grad(T.mean(sqr(-X_hat[0,:,box] + RGB[0,:,box])), Z) via autograd."""
    z = z.clone().requires_grad_(True)
    r1, r2, c1, c2 = int(r1), int(r2), int(c1), int(c2)
    loss = ((-(decode_fn or decode)(P, z)[0, :, r1:r2, c1:c2] + RGB[0, :, r1:r2, c1:c2]) ** 2).mean()
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


# ================================================================================================
# FULL IAN graph (reference IAN.py:67-228; layers.py:207-258, 397-416, 641-853), float32/64 functional
# ================================================================================================

def mdcl(P, name, x, scales):
    """MDCL (layers.py:207-258) through F.conv2d (independent of the tap-loop form in ian_full_numpy)."""
    W = P[name + "W"]
    y = F.conv2d(x, W * P[name + "_coeff_base"].reshape(-1, 1, 1, 1), padding=1)
    for s in scales:
        if s == 0:
            y = y + F.conv2d(x, (W.mean(dim=(2, 3)) * P[name + "_coeff_1x1"].reshape(-1, 1))[:, :, None, None])
        else:
            y = y + F.conv2d(x, W * P[name + "_coeff_%d" % s].reshape(-1, 1, 1, 1), padding=s, dilation=s)
    return y


def mdcl_composite(P, name, x, scales):
    """mdclW formulation (layers.py:138-150 idea): ONE conv with the composite (2*smax+1)^2 kernel.
    KAT: must equal mdcl()."""
    W = P[name + "W"]
    smax = max([1] + [s for s in scales])
    K = 2 * smax + 1
    comp = torch.zeros(W.shape[0], W.shape[1], K, K, dtype=W.dtype)

    def put(s, coeff):
        for i in range(3):
            for j in range(3):
                comp[:, :, smax + (i - 1) * s, smax + (j - 1) * s] += W[:, :, i, j] * coeff.reshape(-1, 1)
    put(1, P[name + "_coeff_base"])
    for s in scales:
        if s == 0:
            comp[:, :, smax, smax] += W.mean(dim=(2, 3)) * P[name + "_coeff_1x1"].reshape(-1, 1)
        else:
            put(s, P[name + "_coeff_%d" % s])
    return F.conv2d(x, comp, padding=smax)


def mdblock(P, name, x, scales, mdcl_fn=mdcl):
    t = _lrelu(_bn(P, name + "bnorm0", x))
    t = _lrelu(_bn(P, name + "bnorm1", mdcl_fn(P, name, t, scales)))
    t = mdcl_fn(P, name + "2", t, scales)
    return _lrelu(_bn(P, name + "bnorm2", x + t))


def made(P, name, z, masks):
    """the MADE layer as wired in the reference graph: its `<name>_input` MaskedLayer runs twice, because
    MADE.__init__ overwrites Layer.input_layer with it (layers.py:769; see oracle/ian_full_numpy.made_forward)."""
    M0, M1, Md = masks
    u = _relu(z @ (P[name + "_input.W"] * M0) + P[name + "_input.b"])
    h = _relu(u @ (P[name + "_input.W"] * M0) + P[name + "_input.b"])
    return h @ (P[name + "_output_W.W"] * M1) + P[name + "_output_W.b"] + u @ (P[name + "_output_D.W"] * Md) + P[name + "_output_D.b"]


def full_encode_mu_ls(P, x):
    h = _lrelu(F.conv2d(x, P["enc_conv1.W"], P["enc_conv1.b"], stride=2, padding=2))
    h = _lrelu(_bn(P, "bnorm2", F.conv2d(h, P["enc_conv2.W"], None, stride=2, padding=2)))
    h = _lrelu(_bn(P, "bnorm3", F.conv2d(h, P["enc_conv3.W"], None, stride=2, padding=2)))
    h = _lrelu(_bn(P, "bnorm4", F.conv2d(h, P["enc_conv4.W"], None, stride=2, padding=2)))
    h = _relu(_bn(P, "bnorm_enc_fc1", h.flatten(1) @ P["enc_fc1.W"]))
    return _bn(P, "mu_bnorm", h @ P["enc_mu.W"]), _bn(P, "ls_bnorm", h @ P["enc_logsigma.W"])


def full_latent(P, z_iaf, masks):
    return (z_iaf - made(P, "l_IAF_mu", z_iaf, masks)) / torch.exp(made(P, "l_IAF_ls", z_iaf, masks))


def full_encode(P, x, masks, deterministic=True, eps=None):
    mu, ls = full_encode_mu_ls(P, x)
    z = mu if deterministic else mu + torch.exp(ls) * eps
    return full_latent(P, z, masks)


def full_decode(P, z, mdcl_fn=mdcl):
    h = _lrelu(z @ P["l_dec_fc2.W"] + P["l_dec_fc2.b"]).reshape(-1, 512, 4, 4)
    h = mdblock(P, "dec_conv2a", deconv(h, P["dec_conv1.W"]), [0, 2], mdcl_fn)
    h = mdblock(P, "dec_conv3a", deconv(h, P["dec_conv2.W"]), [0, 2, 3], mdcl_fn)
    h = mdblock(P, "dec_conv4a", deconv(h, P["dec_conv3.W"]), [0, 2, 3], mdcl_fn)
    h = _lrelu(_bn(P, "bnorm_dc4", deconv(h, P["dec_conv4.W"])))
    sc = [2, 3, 4]
    R = torch.sigmoid(mdcl_fn(P, "R", h, sc))
    G = torch.sigmoid(mdcl_fn(P, "G_a", h, sc) + mdcl_fn(P, "G_b", R, sc))
    B = torch.sigmoid(mdcl_fn(P, "B_a", h, sc) + mdcl_fn(P, "B_b", torch.cat([R, G], 1), sc))

    def beta(a, b):
        return 2.0 * (a / (a + b + 1e-8)) - 1.0
    return torch.stack([beta(R[:, 0], R[:, 1]), beta(G[:, 0], G[:, 1]), beta(B[:, 0], B[:, 1])], 1)


def v1_decode(P, z, mdcl_fn=mdcl):
    """IANv1.py:125-201 (see ian_full_numpy.v1_decode)."""
    h = (z @ P["l_dec_fc2.W"] + P["l_dec_fc2.b"]).reshape(-1, 1024, 4, 4)
    for i, name in ((1, "bnorm_dc1"), (2, "bnorm_dc2"), (3, "bnorm_dc3"), (4, "bnorm_dc4")):
        h = _relu(_bn(P, name, deconv(h, P["dec_conv%d.W" % i])))
    sc = [2, 3, 4]
    R = torch.sigmoid(mdcl_fn(P, "R", h, sc))
    G = torch.sigmoid(mdcl_fn(P, "G_a", h, sc) + mdcl_fn(P, "G_b", R, sc))
    B = torch.sigmoid(mdcl_fn(P, "B_a", h, sc) + mdcl_fn(P, "B_b", torch.cat([R, G], 1), sc))

    def beta(a, b):
        return 2.0 * (a / (a + b + 1e-8)) - 1.0
    return torch.stack([beta(R[:, 0], R[:, 1]), beta(G[:, 0], G[:, 1]), beta(B[:, 0], B[:, 1])], 1)

"""CPU ORACLE (test infrastructure, NOT product code) -- float64 definitional restatement of the FULL IAN
graph (reference IAN.py:67-228) and its custom layers (reference layers.py: MDCL 207-258, beta_layer 397-408,
MDBLOCK 411-416, IAFLayer 641-650, MaskedLayer/DIML/MADE 653-853; mask_generator.py:15-103).

PARITY: pinned to the executed reference (see the header of oracle/ian_numpy.py): IANv1.get_model() and
IAN.get_model() run unmodified on oracle/refshim and this file matches their outputs to 1e-11
(tests/golden/ref_exec_v1.npz, ref_exec_full.npz; tests/test_reference_exec.py).
Third-party semantics restated underneath are the named assumptions of SURVEY.md Appendix C/D; in particular
  C.1  BN(incoming) inside MDBLOCK deletes the preceding DeconvLayer's bias (layers.py:413),
  C.6  DilatedConv2DLayer on a PadLayer(s) input = correlation with taps at (i-1)*s,
  D    MADE masks depend only on `ordering`; ordering = one legacy-RandomState permutation (seed 1234 stream).
"""
from __future__ import annotations

import numpy as np

from . import ian_numpy as on

F64 = np.float64


# ----------------------------------------------------------------------------------------------
# MADE masks (mask_generator.py:15-103 with mask_distribution=0, layers.py:756-763)
# ----------------------------------------------------------------------------------------------

def made_ordering(seed=1234, n=100):
    """ordering after reset("Once") (API.py:33-36 -> layers.py:845-853 -> mask_generator.py:35-38,55-73):
    exactly one shuffle_row_elements draw from RandomStreams(seed).  Recalled Theano seeding (Appendix D):
    RandomStreams keeps gen_seedgen = RandomState(seed); the stream's generator is
    RandomState(gen_seedgen.randint(2**30)); a 1-D shuffle is one permutation(n) of that generator."""
    seedgen = np.random.RandomState(seed)
    stream = np.random.RandomState(int(seedgen.randint(2 ** 30)))
    perm = stream.permutation(n)
    return np.arange(n, dtype=np.float32)[perm]


def made_masks(ordering):
    """_get_mask(in,out) = (conn_in[:,None] <= conn_out[None,:]) (mask_generator.py:93-94); connectivity:
    input = ordering+1, hidden = 1 for every unit (Appendix D), output = ordering.
    Returns float32 0/1 masks (M0 input layer, M1 output layer, Md direct input->output layer)."""
    o = np.asarray(ordering, np.float32)
    conn_in, conn_h, conn_out = o + 1, np.ones_like(o), o
    M0 = (conn_in[:, None] <= conn_h[None, :]).astype(np.float32)       # layers.py:666 (layerIdx 0 -> 1)
    M1 = (conn_h[:, None] <= conn_out[None, :]).astype(np.float32)      # layers.py:790-796 (1 -> 2)
    Md = (conn_in[:, None] <= conn_out[None, :]).astype(np.float32)     # layers.py:696 (_get_mask(0, 2))
    return M0, M1, Md


def made_core(P, name, u, masks):
    """MADE.get_output_for(u) (layers.py:817-818) = get_output(final_layer, {self.z: u}):
    relu(u(W0*M0)+b0)(W1*M1)+b1 + u(Wd*Md)+bd   (MaskedLayer layers.py:666-674, DIML 699-707, ESL at 810)."""
    M0, M1, Md = [np.asarray(m, F64) for m in masks]
    u = np.asarray(u, F64)
    h = made_input_layer(P, name, u, masks)
    out = h @ (np.asarray(P[name + "_output_W.W"], F64) * M1) + np.asarray(P[name + "_output_W.b"], F64)
    return out + u @ (np.asarray(P[name + "_output_D.W"], F64) * Md) + np.asarray(P[name + "_output_D.b"], F64)


def made_input_layer(P, name, z, masks):
    """the MaskedLayer `<name>_input`: relu(z(W0*M0)+b0)  (layers.py:769-775)."""
    M0 = np.asarray(masks[0], F64)
    return on.rectify(np.asarray(z, F64) @ (np.asarray(P[name + "_input.W"], F64) * M0) + np.asarray(P[name + "_input.b"], F64))


def made_forward(P, name, z, masks):
    """What the MADE layer computes INSIDE THE REFERENCE GRAPH.  MADE.__init__ stores its first MaskedLayer in
    `self.input_layer` (layers.py:769) -- the very attribute lasagne.layers.Layer uses to record the layer's input
    (set to `z` by Layer.__init__ at layers.py:738) and that get_all_layers / get_output follow.  So, wired into
    IAN.py:127 / IANv1.py:123, the MADE layer is fed the OUTPUT of `<name>_input` rather than z, and then applies its
    whole stack (input layer included) to that:   made_core(made_input_layer(z)).
    Found by executing the reference's files (tests/golden/make_golden_ref.py); made_core alone is what a reading
    of MADE.get_output_for suggests and is NOT what API.IAN / sample_IAN.py evaluate."""
    return made_core(P, name, made_input_layer(P, name, z, masks), masks)


def iaf(z, mu, ls):
    """IAFLayer (layers.py:641-650)."""
    return (z - mu) / np.exp(ls)


# ----------------------------------------------------------------------------------------------
# MDCL / MDBLOCK / beta
# ----------------------------------------------------------------------------------------------

def _corr3x3_dilated(x, W, s):
    """same-size correlation of x (n,C,H,W) with W (F,C,3,3), taps at (i-1)*s, zero padding."""
    n, c, h, w = x.shape
    xp = np.zeros((n, c, h + 2 * s, w + 2 * s), F64)
    xp[:, :, s:s + h, s:s + w] = x
    y = np.zeros((n, W.shape[0], h, w), F64)
    for i in range(3):
        for j in range(3):
            y += np.einsum("nchw,fc->nfhw", xp[:, :, i * s:i * s + h, j * s:j * s + w], W[:, :, i, j], optimize=True)
    return y


def mdcl(P, name, x, scales):
    """MDCL (layers.py:207-258): base 3x3 + (1x1 mean of W if 0 in scales) + 3x3 dilated s, one shared W (F,C,3,3),
    each branch scaled by a per-output-filter coefficient; summed (ESL), no bias, no nonlinearity."""
    x = np.asarray(x, F64)
    W = np.asarray(P[name + "W"], F64)
    y = _corr3x3_dilated(x, W * np.asarray(P[name + "_coeff_base"], F64)[:, None, None, None], 1)
    for s in scales:
        if s == 0:
            Wm = W.mean(axis=(2, 3)) * np.asarray(P[name + "_coeff_1x1"], F64)[:, None]
            y += np.einsum("nchw,fc->nfhw", x, Wm, optimize=True)
        else:
            y += _corr3x3_dilated(x, W * np.asarray(P[name + "_coeff_%d" % s], F64)[:, None, None, None], s)
    return y


def mdblock(P, name, x, scales):
    """MDBLOCK (layers.py:411-416) with nonlinearity lrelu(0.2) (IAN.py:149,160,171):
    lrelu(BN2(x + MDCL2(lrelu(BN1(MDCL1(lrelu(BN0(x))))))))."""
    t = on.lrelu(on.batchnorm_inf(x, on._bn(P, name + "bnorm0")))
    t = on.lrelu(on.batchnorm_inf(mdcl(P, name, t, scales), on._bn(P, name + "bnorm1")))
    t = mdcl(P, name + "2", t, scales)
    return on.lrelu(on.batchnorm_inf(x + t, on._bn(P, name + "bnorm2")))


def beta(a, b):
    """beta_layer (layers.py:397-408): 2*(alpha/(alpha+beta+1e-8)) - 1."""
    return 2.0 * (a / (a + b + 1e-8)) - 1.0


# ----------------------------------------------------------------------------------------------
# graph (IAN.py:67-228)
# ----------------------------------------------------------------------------------------------

def full_encode_mu_ls(P, x):
    """IAN.py:71-125: as IAN_simple but enc_fc1 uses relu (IAN.py:118)."""
    h = on.lrelu(on.conv5x5_s2(x, P["enc_conv1.W"], P["enc_conv1.b"]))
    h = on.lrelu(on.batchnorm_inf(on.conv5x5_s2(h, P["enc_conv2.W"]), on._bn(P, "bnorm2")))
    h = on.lrelu(on.batchnorm_inf(on.conv5x5_s2(h, P["enc_conv3.W"]), on._bn(P, "bnorm3")))
    h = on.lrelu(on.batchnorm_inf(on.conv5x5_s2(h, P["enc_conv4.W"]), on._bn(P, "bnorm4")))
    h = on.rectify(on.batchnorm_inf(on.dense(h, P["enc_fc1.W"]), on._bn(P, "bnorm_enc_fc1")))
    mu = on.batchnorm_inf(on.dense(h, P["enc_mu.W"]), on._bn(P, "mu_bnorm"))
    ls = on.batchnorm_inf(on.dense(h, P["enc_logsigma.W"]), on._bn(P, "ls_bnorm"))
    return mu, ls


def full_latent(P, z_iaf, masks):
    """l_Z = IAFLayer(l_Z_IAF, MADE_mu(l_Z_IAF), MADE_ls(l_Z_IAF)) (IAN.py:126-128)."""
    return iaf(z_iaf, made_forward(P, "l_IAF_mu", z_iaf, masks), made_forward(P, "l_IAF_ls", z_iaf, masks))


def full_encode(P, x, masks, deterministic=True, eps=None):
    """API.IAN.encode_images for the IAN.py config: get_output(l_Z, deterministic=True) (API.py:50)."""
    mu, ls = full_encode_mu_ls(P, x)
    return full_latent(P, on.gaussian_sample(mu, ls, eps, deterministic), masks)


def full_decode(P, z):
    """IAN.py:129-207 (l_Z -> l_out)."""
    z = np.asarray(z, F64)
    h = on.lrelu(on.dense(z, P["l_dec_fc2.W"], P["l_dec_fc2.b"])).reshape(-1, 512, 4, 4)
    h = on.deconv5x5_s2(h, P["dec_conv1.W"])                    # bias deleted by MDBLOCK's BN(incoming) (C.1)
    h = mdblock(P, "dec_conv2a", h, [0, 2])
    h = on.deconv5x5_s2(h, P["dec_conv2.W"])
    h = mdblock(P, "dec_conv3a", h, [0, 2, 3])
    h = on.deconv5x5_s2(h, P["dec_conv3.W"])
    h = mdblock(P, "dec_conv4a", h, [0, 2, 3])
    h = on.lrelu(on.batchnorm_inf(on.deconv5x5_s2(h, P["dec_conv4.W"]), on._bn(P, "bnorm_dc4")))
    sc = [2, 3, 4]
    R = on.sigmoid(mdcl(P, "R", h, sc))
    G = on.sigmoid(mdcl(P, "G_a", h, sc) + mdcl(P, "G_b", R, sc))
    B = on.sigmoid(mdcl(P, "B_a", h, sc) + mdcl(P, "B_b", np.concatenate([R, G], 1), sc))
    return np.stack([beta(R[:, 0], R[:, 1]), beta(G[:, 0], G[:, 1]), beta(B[:, 0], B[:, 1])], 1)


# ----------------------------------------------------------------------------------------------
# IANv1 graph (reference IANv1.py:63-222): same encoder + MADE/IAF latent as IAN.py; decoder without MDC blocks
# ----------------------------------------------------------------------------------------------

def v1_decode(P, z):
    """IANv1.py:125-201: dense 100->16384 (+bias, linear) -> 4 x [deconv -> BN -> relu] (1024->512->256->128->64)
    -> RGB-Beta head on 64 channels."""
    z = np.asarray(z, F64)
    h = on.dense(z, P["l_dec_fc2.W"], P["l_dec_fc2.b"]).reshape(-1, 1024, 4, 4)       # nonlinearity=None (IANv1.py:127)
    for i, name in ((1, "bnorm_dc1"), (2, "bnorm_dc2"), (3, "bnorm_dc3"), (4, "bnorm_dc4")):
        h = on.rectify(on.batchnorm_inf(on.deconv5x5_s2(h, P["dec_conv%d.W" % i]), on._bn(P, name)))
    sc = [2, 3, 4]
    R = on.sigmoid(mdcl(P, "R", h, sc))
    G = on.sigmoid(mdcl(P, "G_a", h, sc) + mdcl(P, "G_b", R, sc))
    B = on.sigmoid(mdcl(P, "B_a", h, sc) + mdcl(P, "B_b", np.concatenate([R, G], 1), sc))
    return np.stack([beta(R[:, 0], R[:, 1]), beta(G[:, 0], G[:, 1]), beta(B[:, 0], B[:, 1])], 1)

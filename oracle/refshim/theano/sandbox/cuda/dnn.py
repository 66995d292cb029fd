"""theano.sandbox.cuda.dnn stand-in: the cuDNN convolution ops the reference calls, from cuDNN's published definition.

Forward (cudnnConvolutionForward), pad (ph,pw), stride (u,v), filter w[K][C][R][S]:
    CROSS_CORRELATION: y[n,k,p,q] = sum_{c,r,s} x[n,c, p*u + r - ph, q*v + s - pw] * w[k,c,r,s]
    CONVOLUTION      : the same with w[k,c,R-1-r,S-1-s]
GpuDnnConvGradI()(kerns, topgrad, out, desc) = cudnnConvolutionBackwardData: the gradient of that forward w.r.t. x
for dy = topgrad, with x's shape taken from `out`:  dx[n,c,i,j] = sum over (k,p,r), (q,s) with p*u + r - ph = i,
q*v + s - pw = j of dy[n,k,p,q] * w_eff[k,c,r,s].  Written here as a scatter of 25 per-tap matrix products.
"""
import numpy as np

from ... import Var, as_var


class _Desc(object):
    def __init__(self, border_mode, subsample, conv_mode):
        if isinstance(border_mode, int):
            border_mode = (border_mode, border_mode)
        if border_mode == 'valid':
            border_mode = (0, 0)
        assert isinstance(border_mode, tuple), "only explicit padding is used by the reference"
        assert conv_mode in ('conv', 'cross')
        self.pad, self.stride, self.conv_mode = tuple(int(p) for p in border_mode), tuple(int(s) for s in subsample), conv_mode


class GpuDnnConvDesc(object):
    def __init__(self, border_mode, subsample=(1, 1), conv_mode='conv', precision=None):
        self.desc = _Desc(border_mode, subsample, conv_mode)

    def __call__(self, img_shape, kern_shape):
        return self.desc


def conv_forward(x, w, desc):
    (ph, pw), (u, v) = desc.pad, desc.stride
    K, C, R, S = w.shape
    if desc.conv_mode == 'conv':
        w = w[:, :, ::-1, ::-1]
    N, _, H, W = x.shape
    P, Q = (H + 2 * ph - R) // u + 1, (W + 2 * pw - S) // v + 1
    xp = np.zeros((N, C, H + 2 * ph, W + 2 * pw), x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    y = np.zeros((N, K, P, Q), np.result_type(x, w))
    for r in range(R):
        for s in range(S):
            patch = xp[:, :, r:r + u * (P - 1) + 1:u, s:s + v * (Q - 1) + 1:v]          # x[n,c,p*u+r-ph, q*v+s-pw]
            y += np.einsum('ncpq,kc->nkpq', patch, w[:, :, r, s], optimize=True)
    return y


def conv_grad_input(w, dy, out_shape, desc):
    (ph, pw), (u, v) = desc.pad, desc.stride
    K, C, R, S = w.shape
    if desc.conv_mode == 'conv':
        w = w[:, :, ::-1, ::-1]
    N, K2, P, Q = dy.shape
    assert K2 == K, "topgrad channels must equal kerns.shape[0]"
    H, W = int(out_shape[2]), int(out_shape[3])
    assert (H + 2 * ph - R) // u + 1 == P and (W + 2 * pw - S) // v + 1 == Q, "out shape inconsistent with the descriptor"
    buf = np.zeros((N, C, builtin_max(H + 2 * ph, u * (P - 1) + R), builtin_max(W + 2 * pw, v * (Q - 1) + S)), np.result_type(dy, w))
    for r in range(R):
        for s in range(S):
            buf[:, :, r:r + u * (P - 1) + 1:u, s:s + v * (Q - 1) + 1:v] += np.einsum('nkpq,kc->ncpq', dy, w[:, :, r, s], optimize=True)
    return buf[:, :, ph:ph + H, pw:pw + W]


builtin_max = max


class GpuDnnConvGradI(object):
    def __init__(self, inplace=False, workmem=None, algo=None): pass

    def __call__(self, kerns, topgrad, out, desc, alpha=1.0, beta=0.0):
        k, g, o = as_var(kerns), as_var(topgrad), as_var(out)
        return Var(lambda kv, gv, ov: conv_grad_input(kv, gv, ov.shape, desc), [k, g, o], ndim=4)


class GpuDnnConv(object):
    def __init__(self, workmem=None, inplace=False, algo=None): pass

    def __call__(self, img, kerns, out, desc, alpha=1.0, beta=0.0):
        i, k = as_var(img), as_var(kerns)
        return Var(lambda iv, kv: conv_forward(iv, kv, desc), [i, k], ndim=4)


def dnn_conv(img, kerns, border_mode='valid', subsample=(1, 1), conv_mode='conv', direction_hint=None, workmem=None,
             algo=None, precision=None):
    desc = _Desc(border_mode, subsample, conv_mode)
    i, k = as_var(img), as_var(kerns)
    return Var(lambda iv, kv: conv_forward(iv, kv, desc), [i, k], ndim=4)


def dnn_pool(img, ws, stride=(1, 1), mode='max', pad=(0, 0)):
    raise NotImplementedError("pooling is not on the reference's inference path")


def dnn_batch_normalization_train(*a, **k):
    raise NotImplementedError("training-mode batch norm is not on the reference's inference path")

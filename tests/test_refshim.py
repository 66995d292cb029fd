"""The numpy stand-ins for Theano/Lasagne (oracle/refshim) are what lets the reference's own files execute, so their
layer semantics are checked here against an independent library: torch.nn.functional (conv2d is a cross-correlation
by definition; conv_transpose2d is the gradient of conv2d w.r.t. its input; batch_norm in eval mode).  Run in a
subprocess: the stand-ins shadow the names `theano` / `lasagne` / `imp` and must not leak into the pytest process."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BODY = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.nn.functional as F
    import theano, theano.tensor as T, lasagne
    from lasagne.layers import (InputLayer, Conv2DLayer, TransposedConv2DLayer, DilatedConv2DLayer, PadLayer, DenseLayer,
                                SliceLayer, ReshapeLayer, ElemwiseSumLayer, ConcatLayer, NonlinearityLayer, batch_norm,
                                get_output, get_all_params, get_all_layers, get_output_shape)
    from lasagne.layers.dnn import Conv2DDNNLayer
    from theano.sandbox.cuda.dnn import GpuDnnConvDesc, GpuDnnConvGradI
    from theano.sandbox.cuda.basic_ops import gpu_alloc_empty
    rng = np.random.default_rng(0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    X = T.tensor4('X')
    x = rng.standard_normal((2, 3, 12, 12))
    l_in = InputLayer((None, 3, 12, 12))

    def run(layer, **kw):
        return theano.function([X], get_output(layer, X, **kw))(x)

    # forward convolutions: flip_filters=False is a cross-correlation, True a true convolution
    for cls in (Conv2DLayer, Conv2DDNNLayer):
        for flip in (False, True):
            l = cls(l_in, 4, [5, 5], stride=[2, 2], pad=(2, 2), nonlinearity=None, flip_filters=flip, name='c')
            W, b = l.W.get_value(), l.b.get_value()
            ref = F.conv2d(t(x), t(W[:, :, ::-1, ::-1].copy() if flip else W), t(b), stride=2, padding=2).numpy()
            assert np.abs(run(l) - ref).max() < 1e-12, (cls.__name__, flip)
            assert get_output_shape(l) == (None, 4, 6, 6)
    assert Conv2DLayer(l_in, 4, 3).flip_filters is True and Conv2DDNNLayer(l_in, 4, 3).flip_filters is False   # documented defaults

    # transposed convolution = input-gradient of a true convolution (flip_filters=False -> filter_flip=True)
    l = TransposedConv2DLayer(l_in, 4, [5, 5], stride=[2, 2], crop=(1, 1), nonlinearity=None, b=None, name='t')
    W = l.W.get_value()
    assert W.shape == (3, 4, 5, 5) and get_output_shape(l) == (None, 4, 25, 25)
    ref = F.conv_transpose2d(t(x), t(W[:, :, ::-1, ::-1].copy()), stride=2, padding=1).numpy()
    assert np.abs(run(l) - ref).max() < 1e-12

    # cuDNN backward-data as the reference's DeconvLayer calls it (layers.py:467-483): conv mode, pad 2, output 2x
    kern = theano.shared(rng.standard_normal((3, 4, 5, 5)))
    out = gpu_alloc_empty(X.shape[0], kern.shape[1], X.shape[2] * 2, X.shape[3] * 2)
    desc = GpuDnnConvDesc(border_mode=(2, 2), subsample=(2, 2), conv_mode='conv')(out.shape, kern.shape)
    got = theano.function([X], GpuDnnConvGradI()(kern, X, out, desc))(x)
    Wk = kern.get_value()
    ref = F.conv_transpose2d(t(x), t(Wk[:, :, ::-1, ::-1].copy()), stride=2, padding=2, output_padding=1).numpy()
    assert got.shape == (2, 4, 24, 24) and np.abs(got - ref).max() < 1e-12
    # ... which is the autograd input-gradient of the forward true convolution it is defined by
    xi = torch.zeros(2, 4, 24, 24, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xi, t(Wk[:, :, ::-1, ::-1].copy()), stride=2, padding=2)
    (y * t(x)).sum().backward()
    assert np.abs(got - xi.grad.numpy()).max() < 1e-12

    # dilated convolution on a padded input (MDCL, layers.py:246-254): W is (in, out, 3, 3), correlation
    l = DilatedConv2DLayer(PadLayer(l_in, (3, 3)), 4, [3, 3], dilation=(3, 3), nonlinearity=None, b=None, name='d')
    W = l.W.get_value()
    ref = F.conv2d(t(x), t(W.transpose(1, 0, 2, 3).copy()), padding=3, dilation=3).numpy()
    assert W.shape == (3, 4, 3, 3) and np.abs(run(l) - ref).max() < 1e-12

    # batch_norm(): bias removed, nonlinearity moved on top, inference uses (x - mean) * (gamma * inv_std) + beta
    conv = Conv2DLayer(l_in, 4, [3, 3], pad=1, nonlinearity=lasagne.nonlinearities.LeakyRectify(0.2), name='e')
    top = batch_norm(conv, name='bn')
    assert conv.b is None and [p.name for p in get_all_params(top)] == ['e.W', 'bn.beta', 'bn.gamma', 'bn.mean', 'bn.inv_std']
    assert [p.name for p in get_all_params(top, trainable=True)] == ['e.W', 'bn.beta', 'bn.gamma']
    bn = top.input_layer
    vals = {k: rng.uniform(0.5, 1.5, 4) for k in ('beta', 'gamma', 'mean', 'inv_std')}
    for k, v in vals.items():
        getattr(bn, k).set_value(v)
    ref = F.batch_norm(F.conv2d(t(x), t(conv.W.get_value()[:, :, ::-1, ::-1].copy()), padding=1), t(vals['mean']),
                       t(1.0 / vals['inv_std'] ** 2), t(vals['gamma']), t(vals['beta']), False, 0.0, 1e-30)
    ref = F.leaky_relu(ref, 0.2).numpy()
    assert np.abs(run(top, deterministic=True) - ref).max() < 1e-12

    # dense on a 4-d input flattens; reshape with [0]; slice; sum; concat; get_output from an intermediate layer
    d = DenseLayer(l_in, 7, nonlinearity=None, name='f')
    assert np.abs(run(d) - (x.reshape(2, -1) @ d.W.get_value() + d.b.get_value())).max() < 1e-12
    r = ReshapeLayer(d, ([0], 7, 1, 1))
    s = SliceLayer(SliceLayer(l_in, slice(1, None), axis=2), slice(1, None), axis=3)
    assert get_output_shape(s) == (None, 3, 11, 11) and np.array_equal(run(s), x[:, :, 1:, 1:])
    assert np.array_equal(run(ElemwiseSumLayer([l_in, l_in])), 2 * x)
    assert run(ConcatLayer([l_in, l_in])).shape == (2, 6, 12, 12) and run(r).shape == (2, 7, 1, 1)
    Z = T.matrix('Z')
    nl = NonlinearityLayer(d, lasagne.nonlinearities.tanh)
    z = rng.standard_normal((2, 7))
    assert np.array_equal(theano.function([Z], get_output(nl, {d: Z}))(z), np.tanh(z))   # treat_as_input, as API.py:46 does
    assert get_all_layers(nl) == [l_in, d, nl]

    # shared variables are read at call time; updates see the old state; T.grad is a numeric derivative
    w = theano.shared(np.array([1.0, 2.0]), 'w')
    f = theano.function([], w * 2, updates=[(w, w + 1)])
    assert np.array_equal(f(), [2, 4]) and np.array_equal(f(), [4, 6]) and np.array_equal(w.get_value(), [3, 4])
    c = T.mean(T.sqr(T.dot(Z, theano.shared(np.arange(14.).reshape(7, 2)))[0, 0:2]))
    g = theano.function([Z], T.grad(c, Z))(z)
    zt = torch.tensor(z, requires_grad=True)
    ((zt @ torch.arange(14., dtype=torch.float64).reshape(7, 2))[0, 0:2] ** 2).mean().backward()
    assert np.abs(g - zt.grad.numpy()).max() < 1e-6 * np.abs(zt.grad.numpy()).max()
    print("refshim ok")
''') % os.path.join(ROOT, "oracle", "refshim")


def test_standins_match_torch_semantics():
    out = subprocess.run([sys.executable, "-c", BODY], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "refshim ok" in out.stdout, out.stderr[-3000:]

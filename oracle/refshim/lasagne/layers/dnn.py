"""lasagne.layers.dnn stand-in: Conv2DDNNLayer = cuDNN forward convolution, `conv_mode = 'conv' if flip_filters else
'cross'`, border_mode = pad, subsample = stride; note the dnn layer's default is flip_filters=False."""
from .. import init, nonlinearities
from . import BaseConvLayer, Layer, Var, _Desc, _explicit_pad, conv_forward


class Conv2DDNNLayer(BaseConvLayer):
    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), pad=0, untie_biases=False,
                 W=init.GlorotUniform(), b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=False,
                 **kwargs):
        super(Conv2DDNNLayer, self).__init__(incoming, num_filters, filter_size, stride, pad, untie_biases, W, b,
                                             nonlinearity, flip_filters, n=2, **kwargs)

    def convolve(self, input, **kwargs):
        desc = _Desc(_explicit_pad(self.pad, self.filter_size), self.stride, 'conv' if self.flip_filters else 'cross')
        return Var(lambda x, w: conv_forward(x, w, desc), [input, self.W], ndim=4)


class Pool2DDNNLayer(Layer):
    def __init__(self, incoming, pool_size, stride=None, pad=(0, 0), ignore_border=True, mode='max', **kwargs):
        super(Pool2DDNNLayer, self).__init__(incoming, **kwargs)
        raise NotImplementedError("pooling is not on the reference's inference path")


MaxPool2DDNNLayer = Pool2DDNNLayer

"""lasagne.layers stand-in (see ../__init__.py).  Each class follows the Lasagne 0.2.dev1 documentation of the class of
the same name; convolutions evaluate through the cuDNN-definition routines in theano/sandbox/cuda/dnn.py."""
from collections import OrderedDict, deque
from itertools import chain

import numpy as np
import theano
import theano.tensor as T
from theano.sandbox.cuda.dnn import _Desc, conv_forward, conv_grad_input

from .. import init, nonlinearities, utils
from ..utils import as_tuple

Var = theano.Var


# ---- base classes ---------------------------------------------------------------------------------------------
class Layer(object):
    def __init__(self, incoming, name=None):
        if isinstance(incoming, tuple):
            self.input_shape, self.input_layer = incoming, None
        else:
            self.input_shape, self.input_layer = incoming.output_shape, incoming
        self.name = name
        self.params = OrderedDict()
        self.get_output_kwargs = []
        if any(d is not None and d <= 0 for d in self.input_shape):
            raise ValueError("Cannot create Layer with a non-positive input_shape dimension: %r" % (self.input_shape,))

    @property
    def output_shape(self):
        shape = self.get_output_shape_for(self.input_shape)
        if any(isinstance(s, Var) for s in shape):
            raise ValueError("%s returned a symbolic output shape" % type(self).__name__)
        return shape

    def get_params(self, unwrap_shared=True, **tags):
        result = list(self.params.keys())
        only = set(tag for tag, value in tags.items() if value)
        if only:
            result = [p for p in result if not (only - self.params[p])]
        exclude = set(tag for tag, value in tags.items() if not value)
        if exclude:
            result = [p for p in result if not (self.params[p] & exclude)]
        return utils.collect_shared_vars(result) if unwrap_shared else result

    def get_output_shape_for(self, input_shape):
        return input_shape

    def get_output_for(self, input, **kwargs):
        raise NotImplementedError

    def add_param(self, spec, shape, name=None, **tags):
        if name is not None and self.name is not None:
            name = "%s.%s" % (self.name, name)
        param = utils.create_param(spec, shape, name)
        tags['trainable'] = tags.get('trainable', True)
        tags['regularizable'] = tags.get('regularizable', True)
        self.params[param] = set(tag for tag, value in tags.items() if value)
        return param


class MergeLayer(Layer):
    def __init__(self, incomings, name=None):
        self.input_shapes = [i if isinstance(i, tuple) else i.output_shape for i in incomings]
        self.input_layers = [None if isinstance(i, tuple) else i for i in incomings]
        self.name = name
        self.params = OrderedDict()
        self.get_output_kwargs = []

    @Layer.output_shape.getter
    def output_shape(self):
        shape = self.get_output_shape_for(self.input_shapes)
        if any(isinstance(s, Var) for s in shape):
            raise ValueError("%s returned a symbolic output shape" % type(self).__name__)
        return shape


class InputLayer(Layer):
    def __init__(self, shape, input_var=None, name=None, **kwargs):
        self.shape = tuple(shape)
        if input_var is None:
            input_var = T.TensorType(theano.config.floatX, [s == 1 for s in self.shape])(
                "%s.input" % name if name is not None else "input")
        self.input_var, self.name, self.params = input_var, name, OrderedDict()

    @Layer.output_shape.getter
    def output_shape(self):
        return self.shape


# ---- graph helpers --------------------------------------------------------------------------------------------
def get_all_layers(layer, treat_as_input=None):
    queue = deque(layer) if isinstance(layer, (list, tuple)) else deque([layer])
    seen, done, result = set(), set(), []
    if treat_as_input is not None:
        seen.update(treat_as_input)
    while queue:
        layer = queue[0]
        if layer is None:
            queue.popleft()
        elif layer not in seen:
            seen.add(layer)
            if hasattr(layer, 'input_layers'):
                queue.extendleft(reversed(layer.input_layers))
            elif hasattr(layer, 'input_layer'):
                queue.appendleft(layer.input_layer)
        else:
            queue.popleft()
            if layer not in done:
                result.append(layer)
                done.add(layer)
    return result


def get_output(layer_or_layers, inputs=None, **kwargs):
    treat_as_input = list(inputs.keys()) if isinstance(inputs, dict) else []
    all_layers = get_all_layers(layer_or_layers, treat_as_input)
    all_outputs = dict((layer, layer.input_var) for layer in all_layers
                       if isinstance(layer, InputLayer) and layer not in treat_as_input)
    if isinstance(inputs, dict):
        all_outputs.update((layer, utils.as_theano_expression(expr)) for layer, expr in inputs.items())
    elif inputs is not None:
        if len(all_outputs) > 1:
            raise ValueError("get_output() was called with a single input expression on a network with multiple input layers")
        for input_layer in all_outputs:
            all_outputs[input_layer] = utils.as_theano_expression(inputs)
    for layer in all_layers:
        if layer not in all_outputs:
            try:
                if isinstance(layer, MergeLayer):
                    layer_inputs = [all_outputs[l] for l in layer.input_layers]
                else:
                    layer_inputs = all_outputs[layer.input_layer]
            except KeyError:
                raise ValueError("get_output() was called without giving an input expression for the free-floating "
                                 "layer %r" % layer)
            all_outputs[layer] = layer.get_output_for(layer_inputs, **kwargs)
    if isinstance(layer_or_layers, (list, tuple)):
        return [all_outputs[layer] for layer in layer_or_layers]
    return all_outputs[layer_or_layers]


def get_output_shape(layer_or_layers, input_shapes=None):
    if input_shapes is None or input_shapes == []:
        if isinstance(layer_or_layers, (list, tuple)):
            return [l.output_shape for l in layer_or_layers]
        return layer_or_layers.output_shape
    treat_as_input = list(input_shapes.keys()) if isinstance(input_shapes, dict) else []
    all_layers = get_all_layers(layer_or_layers, treat_as_input)
    all_shapes = dict((layer, layer.shape) for layer in all_layers
                      if isinstance(layer, InputLayer) and layer not in treat_as_input)
    if isinstance(input_shapes, dict):
        all_shapes.update(input_shapes)
    else:
        for input_layer in all_shapes:
            all_shapes[input_layer] = input_shapes
    for layer in all_layers:
        if layer not in all_shapes:
            if isinstance(layer, MergeLayer):
                all_shapes[layer] = layer.get_output_shape_for([all_shapes[l] for l in layer.input_layers])
            else:
                all_shapes[layer] = layer.get_output_shape_for(all_shapes[layer.input_layer])
    if isinstance(layer_or_layers, (list, tuple)):
        return [all_shapes[layer] for layer in layer_or_layers]
    return all_shapes[layer_or_layers]


def get_all_params(layer, unwrap_shared=True, **tags):
    layers = get_all_layers(layer)
    params = chain.from_iterable(l.get_params(unwrap_shared=unwrap_shared, **tags) for l in layers)
    return utils.unique(params)


def get_all_param_values(layer, **tags):
    return [p.get_value() for p in get_all_params(layer, **tags)]


def set_all_param_values(layer, values, **tags):
    for p, v in zip(get_all_params(layer, **tags), values):
        p.set_value(v)


# ---- simple layers --------------------------------------------------------------------------------------------
class NonlinearityLayer(Layer):
    def __init__(self, incoming, nonlinearity=nonlinearities.rectify, **kwargs):
        super(NonlinearityLayer, self).__init__(incoming, **kwargs)
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity

    def get_output_for(self, input, **kwargs):
        return self.nonlinearity(input)


class DenseLayer(Layer):
    def __init__(self, incoming, num_units, W=init.GlorotUniform(), b=init.Constant(0.),
                 nonlinearity=nonlinearities.rectify, num_leading_axes=1, **kwargs):
        super(DenseLayer, self).__init__(incoming, **kwargs)
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity
        self.num_units, self.num_leading_axes = num_units, num_leading_axes
        num_inputs = int(np.prod(self.input_shape[num_leading_axes:]))
        self.W = self.add_param(W, (num_inputs, num_units), name="W")
        self.b = None if b is None else self.add_param(b, (num_units,), name="b", regularizable=False)

    def get_output_shape_for(self, input_shape):
        return input_shape[:self.num_leading_axes] + (self.num_units,)

    def get_output_for(self, input, **kwargs):
        if input.ndim > self.num_leading_axes + 1:
            input = input.flatten(self.num_leading_axes + 1)
        activation = T.dot(input, self.W)
        if self.b is not None:
            activation = activation + self.b
        return self.nonlinearity(activation)


class ReshapeLayer(Layer):
    def __init__(self, incoming, shape, **kwargs):
        super(ReshapeLayer, self).__init__(incoming, **kwargs)
        self.shape = tuple(shape)
        self.get_output_shape_for(self.input_shape)

    def get_output_shape_for(self, input_shape, **kwargs):
        out = [input_shape[o[0]] if isinstance(o, list) else o for o in self.shape]
        if -1 in out:
            known = [d for d in out if d != -1]
            total = None if any(d is None for d in input_shape) or any(d is None for d in known) else int(np.prod(input_shape))
            out[out.index(-1)] = None if total is None else total // int(np.prod(known))
        return tuple(out)

    def get_output_for(self, input, **kwargs):
        shape = [input.shape[o[0]] if isinstance(o, list) else o for o in self.shape]
        return input.reshape(tuple(shape))


reshape = ReshapeLayer


class FlattenLayer(Layer):
    def __init__(self, incoming, outdim=2, **kwargs):
        super(FlattenLayer, self).__init__(incoming, **kwargs)
        self.outdim = outdim

    def get_output_shape_for(self, input_shape):
        rest = input_shape[self.outdim - 1:]
        return input_shape[:self.outdim - 1] + (None if any(d is None for d in rest) else int(np.prod(rest)),)

    def get_output_for(self, input, **kwargs):
        return input.flatten(self.outdim)


class SliceLayer(Layer):
    def __init__(self, incoming, indices, axis=-1, **kwargs):
        super(SliceLayer, self).__init__(incoming, **kwargs)
        self.slice, self.axis = indices, axis

    def get_output_shape_for(self, input_shape):
        out = list(input_shape)
        if isinstance(self.slice, int):
            del out[self.axis]
        elif input_shape[self.axis] is not None:
            out[self.axis] = len(range(*self.slice.indices(input_shape[self.axis])))
        return tuple(out)

    def get_output_for(self, input, **kwargs):
        axis = self.axis if self.axis >= 0 else self.axis + input.ndim
        return input[(slice(None),) * axis + (self.slice,)]


class ElemwiseMergeLayer(MergeLayer):
    def __init__(self, incomings, merge_function, cropping=None, **kwargs):
        super(ElemwiseMergeLayer, self).__init__(incomings, **kwargs)
        assert cropping is None
        self.merge_function = merge_function

    def get_output_shape_for(self, input_shapes):
        out = list(input_shapes[0])
        for s in input_shapes[1:]:
            assert len(s) == len(out), "Mismatch: not all input shapes are the same"
            for i, d in enumerate(s):
                if out[i] is None:
                    out[i] = d
                elif d is not None and d != out[i]:
                    raise ValueError("Mismatch: not all input shapes are the same: %r" % (input_shapes,))
        return tuple(out)

    def get_output_for(self, inputs, **kwargs):
        output = None
        for input in inputs:
            output = input if output is None else self.merge_function(output, input)
        return output


class ElemwiseSumLayer(ElemwiseMergeLayer):
    def __init__(self, incomings, coeffs=1, cropping=None, **kwargs):
        super(ElemwiseSumLayer, self).__init__(incomings, T.add, cropping=cropping, **kwargs)
        self.coeffs = [coeffs] * len(incomings) if not isinstance(coeffs, (list, tuple)) else list(coeffs)

    def get_output_for(self, inputs, **kwargs):
        inputs = [i if c == 1 else i * c for c, i in zip(self.coeffs, inputs)]
        return super(ElemwiseSumLayer, self).get_output_for(inputs, **kwargs)


class ConcatLayer(MergeLayer):
    def __init__(self, incomings, axis=1, cropping=None, **kwargs):
        super(ConcatLayer, self).__init__(incomings, **kwargs)
        assert cropping is None
        self.axis = axis

    def get_output_shape_for(self, input_shapes):
        out = list(input_shapes[0])
        sizes = [s[self.axis] for s in input_shapes]
        out[self.axis] = None if any(s is None for s in sizes) else sum(sizes)
        return tuple(out)

    def get_output_for(self, inputs, **kwargs):
        return T.concatenate(inputs, axis=self.axis)


concat = ConcatLayer


class GlobalPoolLayer(Layer):
    def __init__(self, incoming, pool_function=T.mean, **kwargs):
        super(GlobalPoolLayer, self).__init__(incoming, **kwargs)
        self.pool_function = pool_function

    def get_output_shape_for(self, input_shape):
        return input_shape[:2]

    def get_output_for(self, input, **kwargs):
        return self.pool_function(input.flatten(3), axis=2)


class PadLayer(Layer):
    """pads every axis from batch_ndim on; `width`: int, or one entry per padded axis (int = both sides, pair = (l, r))"""
    def __init__(self, incoming, width, val=0, batch_ndim=2, **kwargs):
        super(PadLayer, self).__init__(incoming, **kwargs)
        self.width, self.val, self.batch_ndim = width, val, batch_ndim

    def _widths(self, ndim):
        n = ndim - self.batch_ndim
        w = [self.width] * n if isinstance(self.width, int) else list(self.width)
        assert len(w) == n
        return [(0, 0)] * self.batch_ndim + [(x, x) if isinstance(x, int) else tuple(x) for x in w]

    def get_output_shape_for(self, input_shape):
        return tuple(None if d is None else d + l + r for d, (l, r) in zip(input_shape, self._widths(len(input_shape))))

    def get_output_for(self, input, **kwargs):
        widths, val = self._widths(input.ndim), self.val
        return Var(lambda a: np.pad(a, widths, mode='constant', constant_values=val), [input], ndim=input.ndim)


pad = PadLayer


class Upscale2DLayer(Layer):
    def __init__(self, incoming, scale_factor, mode='repeat', **kwargs):
        super(Upscale2DLayer, self).__init__(incoming, **kwargs)
        self.scale_factor = as_tuple(scale_factor, 2)
        assert mode == 'repeat'

    def get_output_shape_for(self, input_shape):
        a, b = self.scale_factor
        return input_shape[:2] + (None if input_shape[2] is None else input_shape[2] * a,
                                  None if input_shape[3] is None else input_shape[3] * b)

    def get_output_for(self, input, **kwargs):
        a, b = self.scale_factor
        return Var(lambda x: x.repeat(a, axis=2).repeat(b, axis=3), [input], ndim=4)


class ExpressionLayer(Layer):
    def __init__(self, incoming, function, output_shape=None, **kwargs):
        super(ExpressionLayer, self).__init__(incoming, **kwargs)
        self.function, self._shape = function, output_shape

    def get_output_shape_for(self, input_shape):
        return input_shape if self._shape is None else self._shape

    def get_output_for(self, input, **kwargs):
        return self.function(input)


# ---- batch normalisation --------------------------------------------------------------------------------------
class BatchNormLayer(Layer):
    def __init__(self, incoming, axes='auto', epsilon=1e-4, alpha=0.1, beta=init.Constant(0), gamma=init.Constant(1),
                 mean=init.Constant(0), inv_std=init.Constant(1), **kwargs):
        super(BatchNormLayer, self).__init__(incoming, **kwargs)
        if axes == 'auto':
            axes = (0,) + tuple(range(2, len(self.input_shape)))       # all but the second: per channel / per unit
        elif isinstance(axes, int):
            axes = (axes,)
        self.axes, self.epsilon, self.alpha = axes, epsilon, alpha
        shape = [size for axis, size in enumerate(self.input_shape) if axis not in self.axes]
        if any(size is None for size in shape):
            raise ValueError("BatchNormLayer needs specified input sizes for all axes not normalized over.")
        self.beta = None if beta is None else self.add_param(beta, shape, 'beta', trainable=True, regularizable=False)
        self.gamma = None if gamma is None else self.add_param(gamma, shape, 'gamma', trainable=True, regularizable=True)
        self.mean = self.add_param(mean, shape, 'mean', trainable=False, regularizable=False)
        self.inv_std = self.add_param(inv_std, shape, 'inv_std', trainable=False, regularizable=False)

    def get_output_for(self, input, deterministic=False, batch_norm_use_averages=None, batch_norm_update_averages=None, **kwargs):
        use_averages = deterministic if batch_norm_use_averages is None else batch_norm_use_averages
        if use_averages:
            mean, inv_std = self.mean, self.inv_std
        else:                                   # training-mode statistics (never used for the fixtures)
            mean = input.mean(self.axes)
            inv_std = T.inv(T.sqrt(input.var(self.axes) + self.epsilon))
        param_axes = iter(range(input.ndim - len(self.axes)))
        pattern = ['x' if input_axis in self.axes else next(param_axes) for input_axis in range(input.ndim)]
        beta = 0 if self.beta is None else self.beta.dimshuffle(pattern)
        gamma = 1 if self.gamma is None else self.gamma.dimshuffle(pattern)
        mean, inv_std = mean.dimshuffle(pattern), inv_std.dimshuffle(pattern)
        return (input - mean) * (gamma * inv_std) + beta


def batch_norm(layer, **kwargs):
    """lasagne.layers.batch_norm: strip the layer's nonlinearity and bias, insert BatchNormLayer, re-apply the
    nonlinearity in a NonlinearityLayer on top."""
    nonlinearity = getattr(layer, 'nonlinearity', None)
    if nonlinearity is not None:
        layer.nonlinearity = nonlinearities.identity
    if hasattr(layer, 'b') and layer.b is not None:
        del layer.params[layer.b]
        layer.b = None
    bn_name = kwargs.pop('name', None) or (getattr(layer, 'name', None) and layer.name + '_bn')
    layer = BatchNormLayer(layer, name=bn_name, **kwargs)
    if nonlinearity is not None:
        layer = NonlinearityLayer(layer, nonlinearity, name=bn_name and bn_name + '_nonlin')
    return layer


# ---- convolutions ---------------------------------------------------------------------------------------------
def conv_output_length(input_length, filter_size, stride, pad=0):
    if input_length is None:
        return None
    if pad == 'valid':
        output_length = input_length - filter_size + 1
    elif pad == 'full':
        output_length = input_length + filter_size - 1
    elif pad == 'same':
        output_length = input_length
    elif isinstance(pad, int):
        output_length = input_length + 2 * pad - filter_size + 1
    else:
        raise ValueError('Invalid pad: {0}'.format(pad))
    return (output_length + stride - 1) // stride


def conv_input_length(output_length, filter_size, stride, pad=0):
    if output_length is None:
        return None
    pad = {'valid': 0, 'full': filter_size - 1, 'same': filter_size // 2}.get(pad, pad)
    return (output_length - 1) * stride - 2 * pad + filter_size


class BaseConvLayer(Layer):
    def __init__(self, incoming, num_filters, filter_size, stride=1, pad=0, untie_biases=False,
                 W=init.GlorotUniform(), b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=True,
                 n=None, **kwargs):
        super(BaseConvLayer, self).__init__(incoming, **kwargs)
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity
        if n is None:
            n = len(self.input_shape) - 2
        elif n != len(self.input_shape) - 2:
            raise ValueError("Tried to create a %dD convolution layer with input shape %r." % (n, self.input_shape))
        self.n, self.num_filters = n, num_filters
        self.filter_size = as_tuple(filter_size, n, int)
        self.flip_filters = flip_filters
        self.stride = as_tuple(stride, n, int)
        self.untie_biases = untie_biases
        if pad == 'same' and any(s % 2 == 0 for s in self.filter_size):
            raise NotImplementedError('`same` padding requires odd filter size.')
        if pad == 'valid':
            self.pad = as_tuple(0, n)
        elif pad in ('full', 'same'):
            self.pad = pad
        else:
            self.pad = as_tuple(pad, n, int)
        self.W = self.add_param(W, self.get_W_shape(), name="W")
        if b is None:
            self.b = None
        else:
            biases_shape = (num_filters,) + self.output_shape[2:] if self.untie_biases else (num_filters,)
            self.b = self.add_param(b, biases_shape, name="b", regularizable=False)

    def get_W_shape(self):
        return (self.num_filters, self.input_shape[1]) + self.filter_size

    def get_output_shape_for(self, input_shape):
        pad = self.pad if isinstance(self.pad, tuple) else (self.pad,) * self.n
        return (input_shape[0], self.num_filters) + tuple(
            conv_output_length(i, f, s, p) for i, f, s, p in zip(input_shape[2:], self.filter_size, self.stride, pad))

    def get_output_for(self, input, **kwargs):
        conved = self.convolve(input, **kwargs)
        if self.b is None:
            activation = conved
        elif self.untie_biases:
            activation = conved + T.shape_padleft(self.b, 1)
        else:
            activation = conved + self.b.dimshuffle(('x', 0) + ('x',) * self.n)
        return self.nonlinearity(activation)

    def convolve(self, input, **kwargs):
        raise NotImplementedError("BaseConvLayer does not implement the convolve() method.")


def _explicit_pad(pad, filter_size):
    if pad == 'full':
        return tuple(f - 1 for f in filter_size)
    if pad == 'same':
        return tuple(f // 2 for f in filter_size)
    return tuple(pad)


class Conv2DLayer(BaseConvLayer):
    """theano.tensor.nnet.conv2d(border_mode=pad, subsample=stride, filter_flip=flip_filters): flip_filters=True is a
    true convolution, False a cross-correlation; W is (num_filters, num_input_channels, rows, cols)."""
    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), pad=0, untie_biases=False,
                 W=init.GlorotUniform(), b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=True,
                 convolution=None, **kwargs):
        super(Conv2DLayer, self).__init__(incoming, num_filters, filter_size, stride, pad, untie_biases, W, b,
                                          nonlinearity, flip_filters, n=2, **kwargs)

    def convolve(self, input, **kwargs):
        desc = _Desc(_explicit_pad(self.pad, self.filter_size), self.stride, 'conv' if self.flip_filters else 'cross')
        return Var(lambda x, w: conv_forward(x, w, desc), [input, self.W], ndim=4)


class TransposedConv2DLayer(BaseConvLayer):
    """the gradient of Conv2DLayer(stride, pad=crop, flip_filters) w.r.t. its input (AbstractConv2d_gradInputs with
    filter_flip = not flip_filters ... documented as "the backward pass of a convolution"); W is
    (num_input_channels, num_filters, rows, cols); output length = (in - 1)*stride - 2*crop + filter."""
    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), crop=0, untie_biases=False,
                 W=init.GlorotUniform(), b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=False,
                 output_size=None, **kwargs):
        self.output_size = output_size
        super(TransposedConv2DLayer, self).__init__(incoming, num_filters, filter_size, stride, crop, untie_biases, W, b,
                                                    nonlinearity, flip_filters, n=2, **kwargs)
        self.crop = self.pad
        del self.pad

    def get_W_shape(self):
        return (self.input_shape[1], self.num_filters) + self.filter_size

    def get_output_shape_for(self, input_shape):
        if self.output_size is not None:
            return (input_shape[0], self.num_filters) + tuple(self.output_size)
        crop = getattr(self, 'crop', getattr(self, 'pad', None))
        crop = crop if isinstance(crop, tuple) else (crop,) * self.n
        return (input_shape[0], self.num_filters) + tuple(
            conv_input_length(i, f, s, p) for i, f, s, p in zip(input_shape[2:], self.filter_size, self.stride, crop))

    def convolve(self, input, **kwargs):
        crop = _explicit_pad(self.crop, self.filter_size)
        # Lasagne: op = AbstractConv2d_gradInputs(..., filter_flip=not self.flip_filters)
        desc = _Desc(crop, self.stride, 'conv' if not self.flip_filters else 'cross')
        oshape = self.get_output_shape_for

        def run(x, w):
            out = oshape(x.shape)
            return conv_grad_input(w, x, (x.shape[0], w.shape[1]) + tuple(out[2:]), desc)
        return Var(run, [input, self.W], ndim=4)


Deconv2DLayer = TransposedConv2DLayer


class DilatedConv2DLayer(BaseConvLayer):
    """out[n,f,i,j] = sum_{c,r,s} x[n,c, i + r*dh, j + s*dw] * W[c,f,r,s]  (no padding, no filter flipping; note W is
    (num_input_channels, num_filters, rows, cols)); output length = in - (filter - 1)*dilation."""
    def __init__(self, incoming, num_filters, filter_size, dilation=(1, 1), pad=0, untie_biases=False,
                 W=init.GlorotUniform(), b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=False,
                 **kwargs):
        self.dilation = as_tuple(dilation, 2, int)
        super(DilatedConv2DLayer, self).__init__(incoming, num_filters, filter_size, 1, pad, untie_biases, W, b,
                                                 nonlinearity, flip_filters, n=2, **kwargs)
        if self.pad != (0, 0):
            raise NotImplementedError("DilatedConv2DLayer requires pad=0 / (0,0) / 'valid', but got %r." % (pad,))
        if self.flip_filters:
            raise NotImplementedError("DilatedConv2DLayer requires flip_filters=False.")

    def get_W_shape(self):
        return (self.input_shape[1], self.num_filters) + self.filter_size

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], self.num_filters) + tuple(
            conv_output_length(i, (f - 1) * d + 1, 1, 0) for i, f, d in zip(input_shape[2:], self.filter_size, self.dilation))

    def convolve(self, input, **kwargs):
        (dh, dw), (R, S) = self.dilation, self.filter_size

        def run(x, w):
            P, Q = x.shape[2] - (R - 1) * dh, x.shape[3] - (S - 1) * dw
            y = 0.0
            for r in range(R):
                for s in range(S):
                    y = y + np.einsum('ncpq,cf->nfpq', x[:, :, r * dh:r * dh + P, s * dw:s * dw + Q], w[:, :, r, s], optimize=True)
            return y
        return Var(run, [input, self.W], ndim=4)


from . import conv, dnn      # noqa: E402,F401

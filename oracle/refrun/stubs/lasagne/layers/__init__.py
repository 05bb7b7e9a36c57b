"""lasagne.layers stand-in (test infrastructure; see oracle/refrun/__init__.py).

The graph helpers (get_all_layers, get_output, get_all_params, set_all_param_values) follow Lasagne's traversal order
exactly, because that order IS the checkpoint format (np.savez(*get_all_param_values(net))).  The layers compute with torch
CPU float32 tensors inside theano-stand-in Variables.
"""
from collections import OrderedDict, deque
from itertools import chain

import numpy as np
import theano
import theano.tensor as T
import torch
import torch.nn.functional as F
from theano import Variable

from .. import init, nonlinearities
from ..random import get_rng
from ..utils import as_tuple, unique


# ------------------------------------------------------------------ base classes ----
class Layer(object):
    def __init__(self, incoming, name=None):
        if isinstance(incoming, tuple):
            self.input_shape, self.input_layer = incoming, None
        else:
            self.input_shape, self.input_layer = incoming.output_shape, incoming
        self.name = name
        self.params = OrderedDict()
        self.get_output_kwargs = []

    @property
    def output_shape(self):
        shape = self.get_output_shape_for(self.input_shape)
        if any(isinstance(s, Variable) for s in shape):
            raise ValueError('%s returned a symbolic output shape' % self.__class__.__name__)
        return shape

    def get_params(self, unwrap_shared=True, **tags):
        result = list(self.params.keys())
        only = set(tag for tag, value in tags.items() if value)
        if only:
            result = [p for p in result if not (only - self.params[p])]
        exclude = set(tag for tag, value in tags.items() if not value)
        if exclude:
            result = [p for p in result if not (self.params[p] & exclude)]
        return result

    def get_output_shape_for(self, input_shape):
        return input_shape

    def get_output_for(self, input, **kwargs):
        raise NotImplementedError

    def add_param(self, spec, shape, name=None, **tags):
        if name is not None and self.name is not None:
            name = '%s.%s' % (self.name, name)
        shape = tuple(int(s) for s in shape)
        if isinstance(spec, theano.SharedVariable):
            param = spec
            assert param.get_value().shape == shape, 'shared parameter has shape %s, expected %s' % (param.get_value().shape, shape)
        elif isinstance(spec, np.ndarray):
            assert spec.shape == shape
            param = theano.shared(spec.astype(theano.config.floatX), name=name)
        elif callable(spec):
            arr = np.asarray(spec(shape), dtype=theano.config.floatX)
            assert arr.shape == shape, 'initialiser returned shape %s, expected %s' % (arr.shape, shape)
            param = theano.shared(arr, name=name)
        else:
            raise RuntimeError('cannot initialise parameter %s from %r' % (name, spec))
        tags['trainable'] = tags.get('trainable', True)
        tags['regularizable'] = tags.get('regularizable', True)
        self.params[param] = set(tag for tag, value in tags.items() if value)
        return param


class MergeLayer(Layer):
    def __init__(self, incomings, name=None):
        self.input_shapes = [incoming if isinstance(incoming, tuple) else incoming.output_shape for incoming in incomings]
        self.input_layers = [None if isinstance(incoming, tuple) else incoming for incoming in incomings]
        self.name = name
        self.params = OrderedDict()
        self.get_output_kwargs = []

    @Layer.output_shape.getter
    def output_shape(self):
        return self.get_output_shape_for(self.input_shapes)


class InputLayer(Layer):
    def __init__(self, shape, input_var=None, name=None, **kwargs):
        self.shape = tuple(shape)
        if input_var is None:
            input_var = theano.Placeholder(len(shape), theano.config.floatX, name)
        elif input_var.ndim != len(shape):
            raise ValueError('shape has %d dimensions, but variable has %d' % (len(shape), input_var.ndim))
        self.input_var = input_var
        self.name = name
        self.params = OrderedDict()

    @Layer.output_shape.getter
    def output_shape(self):
        return self.shape


# ------------------------------------------------------------------ graph helpers ----
def get_all_layers(layer, treat_as_input=None):
    try:
        queue = deque(layer)
    except TypeError:
        queue = deque([layer])
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


GET_OUTPUT_LOG = []          # top-level get_output calls: which random nodes each one created (see theano.sandbox.rng_mrg)
_DEPTH = [0]


def get_output(layer_or_layers, inputs=None, **kwargs):
    from theano.sandbox import rng_mrg
    top = _DEPTH[0] == 0
    before = rng_mrg.STATE['created']
    _DEPTH[0] += 1
    try:
        return _get_output(layer_or_layers, inputs, **kwargs)
    finally:
        _DEPTH[0] -= 1
        if top:
            GET_OUTPUT_LOG.append({'kwargs': dict(kwargs), 'created': (before, rng_mrg.STATE['created'])})


def _get_output(layer_or_layers, inputs=None, **kwargs):
    treat_as_input = list(inputs.keys()) if isinstance(inputs, dict) else []
    all_layers = get_all_layers(layer_or_layers, treat_as_input)
    all_outputs = dict((layer, layer.input_var) for layer in all_layers if isinstance(layer, InputLayer) and layer not in treat_as_input)
    if isinstance(inputs, dict):
        all_outputs.update((layer, T.as_tensor_variable(expr)) for layer, expr in inputs.items())
    elif inputs is not None:
        for input_layer in all_outputs:
            all_outputs[input_layer] = T.as_tensor_variable(inputs)
    for layer in all_layers:
        if layer not in all_outputs:
            if isinstance(layer, MergeLayer):
                layer_inputs = [all_outputs[l] for l in layer.input_layers]
            else:
                layer_inputs = all_outputs[layer.input_layer]
            all_outputs[layer] = _cse(layer, layer_inputs, kwargs)
    try:
        return [all_outputs[layer] for layer in layer_or_layers]
    except TypeError:
        return all_outputs[layer_or_layers]


_CSE = {}


def _cse(layer, layer_inputs, kwargs):
    """Theano merges identical sub-graphs when it compiles a function; the reference leans on that (every DePool2D calls
    get_output on the contracting path again, layers/mylayers.py:91-93).  The lazy evaluator gets the same effect by
    returning the SAME expression for the same layer applied to the same input expressions with the same keyword arguments.
    Layers that draw random numbers when not deterministic are never merged."""
    random_layer = isinstance(layer, (DropoutLayer, GaussianNoiseLayer)) and not kwargs.get('deterministic', False)
    ins = layer_inputs if isinstance(layer_inputs, list) else [layer_inputs]
    try:
        key = (layer, tuple(id(v) for v in ins), tuple(sorted(kwargs.items())))
        hash(key)
    except TypeError:
        key = None
    if key is None or random_layer:
        return layer.get_output_for(layer_inputs, **kwargs)
    hit = _CSE.get(key)
    if hit is None:
        hit = _CSE[key] = (layer.get_output_for(layer_inputs, **kwargs), ins)          # `ins` kept alive: ids stay unique
    return hit[0]


def get_output_shape(layer_or_layers, input_shapes=None):
    assert input_shapes is None
    try:
        return [layer.output_shape for layer in layer_or_layers]
    except TypeError:
        return layer_or_layers.output_shape


def get_all_params(layer, unwrap_shared=True, **tags):
    layers = get_all_layers(layer)
    return unique(chain.from_iterable(l.get_params(unwrap_shared=unwrap_shared, **tags) for l in layers))


def count_params(layer, **tags):
    return sum(int(np.prod(p.get_value().shape)) for p in get_all_params(layer, **tags))


def get_all_param_values(layer, **tags):
    return [p.get_value() for p in get_all_params(layer, **tags)]


def set_all_param_values(layer, values, **tags):
    params = get_all_params(layer, **tags)
    if len(params) != len(values):
        raise ValueError('mismatch: got %d values to set %d parameters' % (len(values), len(params)))
    for p, v in zip(params, values):
        if p.get_value().shape != v.shape:
            raise ValueError('mismatch: parameter has shape %r but value to set has shape %r' % (p.get_value().shape, v.shape))
        p.set_value(v)


# ------------------------------------------------------------------ shape arithmetic ----
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
    if pad == 'valid':
        pad = 0
    elif pad == 'full':
        pad = filter_size - 1
    elif pad == 'same':
        pad = filter_size // 2
    if not isinstance(pad, int):
        raise ValueError('Invalid pad: {0}'.format(pad))
    return (output_length - 1) * stride - 2 * pad + filter_size


def pool_output_length(input_length, pool_size, stride, pad, ignore_border):
    if input_length is None or pool_size is None:
        return None
    if ignore_border:
        output_length = input_length + 2 * pad - pool_size + 1
        output_length = (output_length + stride - 1) // stride
    else:
        assert pad == 0
        if stride >= pool_size:
            output_length = (input_length + stride - 1) // stride
        else:
            output_length = max(0, (input_length - pool_size + stride - 1) // stride) + 1
    return output_length


# ------------------------------------------------------------------ convolutions ----
def _int_pad(pad, filter_size):
    """Lasagne's `pad` argument as explicit integers per axis."""
    if pad == 'valid':
        return (0,) * len(filter_size)
    if pad == 'same':
        if any(s % 2 == 0 for s in filter_size):
            raise NotImplementedError('`same` padding requires odd filter size.')
        return tuple(s // 2 for s in filter_size)
    if pad == 'full':
        return tuple(s - 1 for s in filter_size)
    return as_tuple(pad, len(filter_size), int)


class Conv2DLayer(Layer):
    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), pad=0, untie_biases=False, W=init.GlorotUniform(),
                 b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=True, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        assert not untie_biases
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity
        self.num_filters = num_filters
        self.filter_size = as_tuple(filter_size, 2, int)
        self.stride = as_tuple(stride, 2, int)
        self.flip_filters = flip_filters
        if pad in ('valid', 'same', 'full'):
            self.pad = pad if pad != 'valid' else (0, 0)
        else:
            self.pad = as_tuple(pad, 2, int)
        self.W = self.add_param(W, self.get_W_shape(), name='W')
        self.b = None if b is None else self.add_param(b, (num_filters,), name='b', regularizable=False)

    def get_W_shape(self):
        return (self.num_filters, self.input_shape[1]) + self.filter_size

    def get_output_shape_for(self, input_shape):
        pad = self.pad if isinstance(self.pad, tuple) else (self.pad,) * 2
        return (input_shape[0], self.num_filters) + tuple(conv_output_length(i, f, s, p) for i, f, s, p in
                                                          zip(input_shape[2:], self.filter_size, self.stride, pad))

    def convolve(self, input):
        pad, stride, flip = _int_pad(self.pad, self.filter_size), self.stride, self.flip_filters
        return Variable(lambda x, W: F.conv2d(x, W.flip(2, 3) if flip else W, None, stride=stride, padding=pad), [input, self.W], ndim=4)

    def get_output_for(self, input, **kwargs):
        conved = self.convolve(input)
        activation = conved if self.b is None else conved + self.b.dimshuffle('x', 0, 'x', 'x')
        return self.nonlinearity(activation)


class TransposedConv2DLayer(Conv2DLayer):
    """The input-gradient of a forward convolution with `filter_flip = not flip_filters` (Lasagne: "implemented as the
    backward pass of a corresponding non-transposed convolution"); W is (input channels, num_filters, rows, cols)."""

    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), crop=0, untie_biases=False, W=init.GlorotUniform(),
                 b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=False, output_size=None, **kwargs):
        assert output_size is None
        Conv2DLayer.__init__(self, incoming, num_filters, filter_size, stride, crop, untie_biases, W, b, nonlinearity,
                             flip_filters, **kwargs)
        self.crop = self.pad
        del self.pad

    def get_W_shape(self):
        return (self.input_shape[1], self.num_filters) + self.filter_size

    def get_output_shape_for(self, input_shape):
        crop = self.crop if isinstance(self.crop, tuple) else (self.crop,) * 2
        return (input_shape[0], self.num_filters) + tuple(conv_input_length(i, f, s, p) for i, f, s, p in
                                                          zip(input_shape[2:], self.filter_size, self.stride, crop))

    def convolve(self, input):
        crop, stride = _int_pad(self.crop, self.filter_size), self.stride
        forward_flips = not self.flip_filters          # filter_flip of the forward convolution being transposed

        def run(x, W):
            out_hw = tuple(conv_input_length(i, f, s, p) for i, f, s, p in zip(x.shape[2:], self.filter_size, stride, crop))
            kern = W.flip(2, 3) if forward_flips else W          # the forward op correlates with this kernel
            return torch.nn.grad.conv2d_input((x.shape[0], self.num_filters) + out_hw, kern, x, stride=stride, padding=crop)
        return Variable(run, [input, self.W], ndim=4)


Deconv2DLayer = TransposedConv2DLayer


class DilatedConv2DLayer(Conv2DLayer):
    """Lasagne computes the dilated convolution as the weight-gradient of a strided convolution (subsample = dilation)
    between the batch/channel-transposed input and W, W being (input channels, num_filters, rows, cols)."""

    def __init__(self, incoming, num_filters, filter_size, dilation=(1, 1), pad=0, untie_biases=False, W=init.GlorotUniform(),
                 b=init.Constant(0.), nonlinearity=nonlinearities.rectify, flip_filters=False, **kwargs):
        self.dilation = as_tuple(dilation, 2, int)
        Conv2DLayer.__init__(self, incoming, num_filters, filter_size, 1, pad, untie_biases, W, b, nonlinearity, flip_filters, **kwargs)
        if self.pad != (0, 0):
            raise NotImplementedError('DilatedConv2DLayer requires pad=0 / (0,0) / "valid"')
        if self.flip_filters:
            raise NotImplementedError('DilatedConv2DLayer requires flip_filters=False')

    def get_W_shape(self):
        return (self.input_shape[1], self.num_filters) + self.filter_size

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], self.num_filters) + tuple(conv_output_length(i, (f - 1) * d + 1, 1, 0) for i, f, d in
                                                          zip(input_shape[2:], self.filter_size, self.dilation))

    def convolve(self, input):
        def run(x, W):
            out_hw = tuple(i - (f - 1) * d for i, f, d in zip(x.shape[2:], self.filter_size, self.dilation))
            img = x.permute(1, 0, 2, 3).contiguous()           # (channels, batch, rows, cols): channels play the batch
            gw = torch.nn.grad.conv2d_weight(img, (self.num_filters, x.shape[0]) + out_hw, W.contiguous(), stride=self.dilation)
            return gw.permute(1, 0, 2, 3)
        return Variable(run, [input, self.W], ndim=4)


# ------------------------------------------------------------------ pooling ----
class _MaxPoolTheano(torch.autograd.Function):
    """Max pooling, ignore_border=True, with Theano's CPU MaxPoolGrad: every element EQUAL to its window's maximum receives
    that window's gradient (`if maxout == x: gx += gz`), not only the first one."""

    @staticmethod
    def forward(ctx, x, ws, st):
        assert ws == st, 'the reference only pools with stride = pool size'
        out = F.max_pool2d(x, ws, st)
        ctx.save_for_backward(x, out)
        ctx.ws = ws
        return out

    @staticmethod
    def backward(ctx, gz):
        x, out = ctx.saved_tensors
        a, b = ctx.ws
        gx = torch.zeros_like(x)
        for i in range(out.shape[2]):
            for r in range(a):
                row = x[:, :, i * a + r, :out.shape[3] * b].reshape(x.shape[0], x.shape[1], out.shape[3], b)
                hit = (row == out[:, :, i, :, None]).to(gz.dtype) * gz[:, :, i, :, None]
                gx[:, :, i * a + r, :out.shape[3] * b] = hit.reshape(x.shape[0], x.shape[1], -1)
        return gx, None, None


class Pool2DLayer(Layer):
    def __init__(self, incoming, pool_size, stride=None, pad=(0, 0), ignore_border=True, mode='max', **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.pool_size = as_tuple(pool_size, 2)
        self.stride = self.pool_size if stride is None else as_tuple(stride, 2)
        self.pad = as_tuple(pad, 2)
        self.ignore_border, self.mode = ignore_border, mode
        assert mode == 'max' and ignore_border and self.pad == (0, 0)

    def get_output_shape_for(self, input_shape):
        output_shape = list(input_shape)
        for d in (0, 1):
            output_shape[2 + d] = pool_output_length(input_shape[2 + d], self.pool_size[d], self.stride[d], self.pad[d], self.ignore_border)
        return tuple(output_shape)

    def get_output_for(self, input, **kwargs):
        return Variable(lambda x: _MaxPoolTheano.apply(x, self.pool_size, self.stride), [input], ndim=4)


class MaxPool2DLayer(Pool2DLayer):
    pass


class Upscale2DLayer(Layer):
    def __init__(self, incoming, scale_factor, mode='repeat', **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.scale_factor = as_tuple(scale_factor, 2)

    def get_output_shape_for(self, input_shape):
        output_shape = list(input_shape)
        if output_shape[2] is not None:
            output_shape[2] *= self.scale_factor[0]
        if output_shape[3] is not None:
            output_shape[3] *= self.scale_factor[1]
        return tuple(output_shape)

    def get_output_for(self, input, **kwargs):
        a, b = self.scale_factor
        upscaled = input
        if b > 1:
            upscaled = T.extra_ops.repeat(upscaled, b, 3)
        if a > 1:
            upscaled = T.extra_ops.repeat(upscaled, a, 2)
        return upscaled


class InverseLayer(MergeLayer):
    def __init__(self, incoming, layer, **kwargs):
        MergeLayer.__init__(self, [incoming, layer, getattr(layer, 'input_layer', None) or getattr(layer, 'input_layers', None)], **kwargs)

    def get_output_shape_for(self, input_shapes):
        return input_shapes[2]

    def get_output_for(self, inputs, **kwargs):
        input, layer_out, layer_in = inputs
        return theano.grad(None, wrt=layer_in, known_grads={layer_out: input})


# ------------------------------------------------------------------ merging ----
def autocrop(inputs, cropping):
    if cropping is None:
        return inputs
    ndim = inputs[0].ndim
    if not all(input.ndim == ndim for input in inputs):
        raise ValueError('Not all inputs are of the same dimensionality.')
    cropping = list(cropping)
    if ndim > len(cropping):
        cropping = cropping + [None] * (ndim - len(cropping))

    def run(ts, i):
        shapes = [t.shape for t in ts]
        min_shape = [min(s[d] for s in shapes) for d in range(ndim)]
        slices = []
        for dim, cr in enumerate(cropping):
            if cr is None:
                slices.append(slice(None))
            else:
                sz = min_shape[dim]
                if cr == 'lower':
                    slices.append(slice(None, sz))
                elif cr == 'upper':
                    slices.append(slice(shapes[i][dim] - sz, None))
                elif cr == 'center':
                    offset = (shapes[i][dim] - sz) // 2
                    slices.append(slice(offset, offset + sz))
                else:
                    raise ValueError('Unknown crop mode {0!r}'.format(cr))
        return ts[i][tuple(slices)]
    return [Variable(lambda ts, i=i: run(ts, i), [list(inputs)], ndim=ndim) for i in range(len(inputs))]


def autocrop_array_shapes(input_shapes, cropping):
    if cropping is None:
        return input_shapes
    ndim = len(input_shapes[0])
    if not all(len(sh) == ndim for sh in input_shapes):
        raise ValueError('Not all inputs are of the same dimensionality.')
    result = []
    cropping = list(cropping)
    if ndim > len(cropping):
        cropping = cropping + [None] * (ndim - len(cropping))
    for sh, cr in zip(zip(*input_shapes), cropping):
        if cr is None:
            result.append(sh)
        elif cr in ('lower', 'center', 'upper'):
            min_sh = None if any(x is None for x in sh) else min(sh)
            result.append([min_sh] * len(sh))
        else:
            raise ValueError('Unknown crop mode {0!r}'.format(cr))
    return [tuple(sh) for sh in zip(*result)]


class ConcatLayer(MergeLayer):
    def __init__(self, incomings, axis=1, cropping=None, **kwargs):
        MergeLayer.__init__(self, incomings, **kwargs)
        self.axis = axis
        if cropping is not None:
            cropping = list(cropping)
            cropping[axis] = None
        self.cropping = cropping

    def get_output_shape_for(self, input_shapes):
        input_shapes = autocrop_array_shapes(input_shapes, self.cropping)
        output_shape = [next((s for s in sizes if s is not None), None) for sizes in zip(*input_shapes)]
        sizes = [input_shape[self.axis] for input_shape in input_shapes]
        output_shape[self.axis] = None if any(s is None for s in sizes) else sum(sizes)
        return tuple(output_shape)

    def get_output_for(self, inputs, **kwargs):
        inputs = autocrop(inputs, self.cropping)
        return T.concatenate(inputs, axis=self.axis)


class ElemwiseMergeLayer(MergeLayer):
    def __init__(self, incomings, merge_function, cropping=None, **kwargs):
        MergeLayer.__init__(self, incomings, **kwargs)
        self.merge_function = merge_function
        self.cropping = cropping

    def get_output_shape_for(self, input_shapes):
        input_shapes = autocrop_array_shapes(input_shapes, self.cropping)
        return tuple(next((s for s in sizes if s is not None), None) for sizes in zip(*input_shapes))

    def get_output_for(self, inputs, **kwargs):
        inputs = autocrop(inputs, self.cropping)
        output = None
        for input in inputs:
            output = self.merge_function(output, input) if output is not None else input
        return output


class ElemwiseSumLayer(ElemwiseMergeLayer):
    def __init__(self, incomings, coeffs=1, cropping=None, **kwargs):
        ElemwiseMergeLayer.__init__(self, incomings, T.add, cropping=cropping, **kwargs)
        self.coeffs = [coeffs] * len(incomings) if not isinstance(coeffs, list) else coeffs

    def get_output_for(self, inputs, **kwargs):
        inputs = [input * coeff if coeff != 1 else input for coeff, input in zip(self.coeffs, inputs)]
        return ElemwiseMergeLayer.get_output_for(self, inputs, **kwargs)


# ------------------------------------------------------------------ shape layers ----
class NonlinearityLayer(Layer):
    def __init__(self, incoming, nonlinearity=nonlinearities.rectify, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity

    def get_output_for(self, input, **kwargs):
        return self.nonlinearity(input)


class DimshuffleLayer(Layer):
    def __init__(self, incoming, pattern, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.pattern = pattern

    def get_output_shape_for(self, input_shape):
        return tuple(1 if p == 'x' else input_shape[p] for p in self.pattern)

    def get_output_for(self, input, **kwargs):
        return input.dimshuffle(self.pattern)


class ReshapeLayer(Layer):
    def __init__(self, incoming, shape, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        shape = tuple(shape)
        for s in shape:
            if isinstance(s, int):
                if s == 0 or s < -1:
                    raise ValueError('`shape` integers must be positive or -1')
            elif isinstance(s, list):
                if len(s) != 1 or not isinstance(s[0], int) or s[0] < 0:
                    raise ValueError('`shape` input references must be single-element lists of int >= 0')
            elif isinstance(s, Variable):
                if s.ndim != 0:
                    raise ValueError('A symbolic variable in a shape specification must be a scalar, but had %i dimensions' % s.ndim)
            else:
                raise ValueError('`shape` must be a tuple of int and/or [int]')
        self.shape = shape

    def get_output_shape_for(self, input_shape, **kwargs):
        output_shape = list(self.shape)
        masked_input_shape = list(input_shape)
        for dim, o in enumerate(output_shape):
            if isinstance(o, list):
                output_shape[dim] = input_shape[o[0]]
                masked_input_shape[o[0]] = 1
            elif isinstance(o, Variable):
                output_shape[dim] = None
        if -1 in output_shape and not any(s is None for s in output_shape + masked_input_shape):
            known = int(np.prod([s for s in output_shape if s != -1]))
            output_shape[output_shape.index(-1)] = int(np.prod(input_shape)) // known
        elif -1 in output_shape:
            output_shape[output_shape.index(-1)] = None
        return tuple(output_shape)

    def get_output_for(self, input, **kwargs):
        output_shape = list(self.shape)
        for dim, o in enumerate(output_shape):
            if isinstance(o, list):
                output_shape[dim] = input.shape[o[0]]
        return input.reshape(tuple(output_shape))


class PadLayer(Layer):
    def __init__(self, incoming, width, val=0, batch_ndim=2, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.width, self.val, self.batch_ndim = width, val, batch_ndim
        assert isinstance(width, int)

    def get_output_shape_for(self, input_shape):
        return tuple(s if (k < self.batch_ndim or s is None) else s + 2 * self.width for k, s in enumerate(input_shape))

    def get_output_for(self, input, **kwargs):
        w, val, bn = self.width, self.val, self.batch_ndim

        def run(x):
            out = torch.full([s if k < bn else s + 2 * w for k, s in enumerate(x.shape)], float(val), dtype=x.dtype)
            out[(slice(None),) * bn + (slice(w, -w),) * (x.dim() - bn)] = x
            return out
        return Variable(run, [input], ndim=input._ndim)


# ------------------------------------------------------------------ noise / normalisation ----
class DropoutLayer(Layer):
    def __init__(self, incoming, p=0.5, rescale=True, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        self.p, self.rescale = p, rescale

    def get_output_for(self, input, deterministic=False, **kwargs):
        if deterministic or self.p == 0:
            return input
        # The reference builds one non-deterministic graph per net only to read its symbolic SHAPE (models/fcn8.py:124,
        # models/fcn_up.py:156); Theano infers that shape without running the dropout.  Here the graph is evaluated, with a
        # torch mask instead of the MRG stream: the values are never consumed, only the shape is.
        q = 1.0 - self.p

        def run(x):
            m = (torch.rand(x.shape) < q).to(x.dtype)
            return x * m / q if self.rescale else x * m
        return Variable(run, [input], ndim=input._ndim)


class GaussianNoiseLayer(Layer):
    def __init__(self, incoming, sigma=0.1, **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        from theano.sandbox.rng_mrg import MRG_RandomStreams as RandomStreams
        self._srng = RandomStreams(get_rng().randint(1, 2147462579))
        self.sigma = sigma

    def get_output_for(self, input, deterministic=False, **kwargs):
        if deterministic or self.sigma == 0:
            return input
        return input + self._srng.normal(input.shape, avg=0.0, std=self.sigma)          # a NEW random node per symbolic call


class BatchNormLayer(Layer):
    def __init__(self, incoming, axes='auto', epsilon=1e-4, alpha=0.1, beta=init.Constant(0), gamma=init.Constant(1),
                 mean=init.Constant(0), inv_std=init.Constant(1), **kwargs):
        Layer.__init__(self, incoming, **kwargs)
        if axes == 'auto':
            axes = (0,) + tuple(range(2, len(self.input_shape)))
        elif isinstance(axes, int):
            axes = (axes,)
        self.axes, self.epsilon, self.alpha = axes, epsilon, alpha
        shape = [size for axis, size in enumerate(self.input_shape) if axis not in self.axes]
        if any(size is None for size in shape):
            raise ValueError('BatchNormLayer needs specified input sizes for all axes not normalized over.')
        self.beta = None if beta is None else self.add_param(beta, shape, 'beta', trainable=True, regularizable=False)
        self.gamma = None if gamma is None else self.add_param(gamma, shape, 'gamma', trainable=True, regularizable=True)
        self.mean = self.add_param(mean, shape, 'mean', trainable=False, regularizable=False)
        self.inv_std = self.add_param(inv_std, shape, 'inv_std', trainable=False, regularizable=False)

    def get_output_for(self, input, deterministic=False, batch_norm_use_averages=None, batch_norm_update_averages=None, **kwargs):
        input_mean = input.mean(self.axes)
        input_inv_std = T.inv(T.sqrt(input.var(self.axes) + self.epsilon))
        use_averages = deterministic if batch_norm_use_averages is None else batch_norm_use_averages
        mean, inv_std = (self.mean, self.inv_std) if use_averages else (input_mean, input_inv_std)
        param_axes = iter(range(input.ndim - len(self.axes)))
        pattern = ['x' if input_axis in self.axes else next(param_axes) for input_axis in range(input.ndim)]
        beta = 0 if self.beta is None else self.beta.dimshuffle(pattern)
        gamma = 1 if self.gamma is None else self.gamma.dimshuffle(pattern)
        mean = mean.dimshuffle(pattern)
        inv_std = inv_std.dimshuffle(pattern)
        return (input - mean) * (gamma * inv_std) + beta


def batch_norm(layer, **kwargs):
    nonlinearity = getattr(layer, 'nonlinearity', None)
    if nonlinearity is not None:
        layer.nonlinearity = nonlinearities.identity
    if hasattr(layer, 'b') and layer.b is not None:
        del layer.params[layer.b]
        layer.b = None
    layer = BatchNormLayer(layer, **kwargs)
    if nonlinearity is not None:
        layer = NonlinearityLayer(layer, nonlinearity)
    return layer


from . import merge, pool  # noqa: E402,F401

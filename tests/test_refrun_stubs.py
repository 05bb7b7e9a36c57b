"""The harness that executes the reference (oracle/refrun) is itself test infrastructure that parity rests on, so its pieces
are checked here against definitions written out as explicit loops -- Lasagne's / Theano's documented layer semantics on tiny
tensors -- and the Python-2 source rewriting against known snippets.  Needs neither a GPU nor /root/reference."""
import itertools

import numpy as np
import pytest

from oracle.refrun import py2import

py2import.install_stubs()
import lasagne  # noqa: E402
import theano  # noqa: E402
import theano.tensor as T  # noqa: E402
from lasagne.layers import (BatchNormLayer, ConcatLayer, Conv2DLayer, Deconv2DLayer, DilatedConv2DLayer, ElemwiseSumLayer,  # noqa: E402
                            InputLayer, InverseLayer, PadLayer, Pool2DLayer, get_all_layers, get_all_params, get_output)
from lasagne.nonlinearities import linear  # noqa: E402

RNG = np.random.RandomState(7)


def _run(layer, x, **kw):
    xv = T.tensor4('x')
    lin = InputLayer((None, x.shape[1], None, None), xv)
    out = layer(lin)
    return out, theano.function([xv], get_output(out, **kw))(x)


def test_conv2d_layer_is_correlation_or_convolution_by_definition():
    x = RNG.randn(2, 3, 6, 7).astype(np.float32)
    for flip, pad in itertools.product((False, True), (0, 1, 'same')):
        lay, got = _run(lambda l: Conv2DLayer(l, 4, 3, pad=pad, flip_filters=flip, nonlinearity=linear), x)
        W, b = lay.W.get_value(), RNG.randn(4).astype(np.float32)
        lay.b.set_value(b)
        got = theano.function([lay.input_layer.input_var], get_output(lay))(x)
        p = 1 if pad == 'same' else pad
        xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
        want = np.zeros((2, 4, xp.shape[2] - 2, xp.shape[3] - 2))
        for n, f, i, j in np.ndindex(*want.shape):
            for c, r, s in np.ndindex(3, 3, 3):
                w = W[f, c, 2 - r, 2 - s] if flip else W[f, c, r, s]          # flip_filters=True: a true convolution
                want[n, f, i, j] += w * xp[n, c, i + r, j + s]
            want[n, f, i, j] += b[f]
        assert np.abs(got - want).max() < 1e-5


def test_deconv2d_layer_is_the_input_gradient_of_a_true_convolution():
    """Deconv2DLayer(flip_filters=False, crop='valid'): out[n, f, s*i + r, s*j + q] += x[n, c, i, j] * W[c, f, k-1-r, k-1-q]."""
    x = RNG.randn(2, 3, 4, 5).astype(np.float32)
    for k, s in ((4, 2), (3, 2), (2, 1)):
        lay, got = _run(lambda l: Deconv2DLayer(l, 2, k, stride=s, crop='valid', nonlinearity=linear), x)
        W = lay.W.get_value()
        assert W.shape == (3, 2, k, k)                                          # (input channels, num_filters, rows, cols)
        want = np.zeros((2, 2, (4 - 1) * s + k, (5 - 1) * s + k))
        for n, c, i, j in np.ndindex(2, 3, 4, 5):
            for f, r, q in np.ndindex(2, k, k):
                want[n, f, s * i + r, s * j + q] += x[n, c, i, j] * W[c, f, k - 1 - r, k - 1 - q]
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-5


def test_dilated_conv_layer_by_definition():
    """DilatedConv2DLayer: W is (input channels, num_filters, k, k); out[n,f,i,j] = b[f] + sum W[c,f,r,s] x[n,c,i+r*d,j+s*d]."""
    x = RNG.randn(2, 3, 11, 12).astype(np.float32)
    for d in (1, 2, 4):
        lay, got = _run(lambda l: DilatedConv2DLayer(l, 5, 3, d, nonlinearity=linear), x)
        W = lay.W.get_value()
        assert W.shape == (3, 5, 3, 3)
        want = np.zeros((2, 5, 11 - 2 * d, 12 - 2 * d))
        for n, f, i, j in np.ndindex(*want.shape):
            for c, r, s in np.ndindex(3, 3, 3):
                want[n, f, i, j] += W[c, f, r, s] * x[n, c, i + r * d, j + s * d]
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-5


def test_pool_ignores_the_border_and_its_gradient_follows_theanos_maxpoolgrad():
    """Pool2DLayer(2): floor((size - 2) / 2) + 1 outputs; Theano's CPU MaxPoolGrad.perform: `if maxout == x: gx += gz` for
    every element of the window (ties all receive the gradient); rows / columns beyond the last window receive nothing."""
    x = RNG.randint(0, 3, size=(2, 2, 5, 7)).astype(np.float32)                  # small integers: plenty of ties
    xv = T.tensor4('x')
    lin = InputLayer((None, 2, None, None), xv)
    pool = Pool2DLayer(lin, 2)
    inp, out = get_output([lin, pool])
    gz = RNG.randn(2, 2, 2, 3).astype(np.float32)
    gzv = T.tensor4('gz')
    g = theano.grad(None, wrt=inp, known_grads={out: gzv})
    got_out, got_g = theano.function([xv, gzv], [out, g])(x, gz)
    want_out = np.zeros((2, 2, 2, 3), np.float32)
    want_g = np.zeros_like(x)
    for n, c, i, j in np.ndindex(2, 2, 2, 3):
        win = x[n, c, 2 * i:2 * i + 2, 2 * j:2 * j + 2]
        want_out[n, c, i, j] = win.max()
        for r, s in np.ndindex(2, 2):
            if win[r, s] == win.max():
                want_g[n, c, 2 * i + r, 2 * j + s] += gz[n, c, i, j]
    assert np.array_equal(got_out, want_out) and np.array_equal(got_g, want_g)
    inv = InverseLayer(InputLayer((None, 2, None, None), gzv), pool)             # lasagne: the same gradient expression
    assert np.array_equal(theano.function([xv, gzv], get_output(inv))(x, gz), want_g)


def test_batchnorm_deterministic_and_batch_statistics():
    x = RNG.randn(3, 4, 5, 6).astype(np.float32)
    xv = T.tensor4('x')
    bn = BatchNormLayer(InputLayer((None, 4, None, None), xv))
    assert [p.name for p in bn.get_params()] == ['beta', 'gamma', 'mean', 'inv_std']
    vals = [RNG.randn(4).astype(np.float32) for _ in range(4)]
    for p, v in zip(bn.get_params(), vals):
        p.set_value(v)
    beta, gamma, mean, inv_std = [v[None, :, None, None] for v in vals]
    det = theano.function([xv], get_output(bn, deterministic=True))(x)
    assert np.abs(det - ((x - mean) * (gamma * inv_std) + beta)).max() < 1e-5
    m = x.mean(axis=(0, 2, 3), keepdims=True)
    istd = 1.0 / np.sqrt(x.var(axis=(0, 2, 3), keepdims=True) + 1e-4)           # biased variance, epsilon 1e-4
    for kw in ({}, {'deterministic': True, 'batch_norm_use_averages': False}):
        got = theano.function([xv], get_output(bn, **kw))(x)
        assert np.abs(got - ((x - m) * (gamma * istd) + beta)).max() < 1e-4
    assert [sorted(bn.params[p]) for p in bn.get_params()] == [['trainable'], ['regularizable', 'trainable'], [], []]
    assert len(bn.get_params(trainable=True)) == 2


def test_centre_cropping_merge_and_concat_order():
    a, b = RNG.randn(1, 2, 9, 10).astype(np.float32), RNG.randn(1, 2, 6, 5).astype(np.float32)
    av, bv = T.tensor4('a'), T.tensor4('b')
    la, lb = InputLayer((None, 2, None, None), av), InputLayer((None, 2, None, None), bv)
    s = ElemwiseSumLayer((la, lb), cropping=[None, None, 'center', 'center'])
    got = theano.function([av, bv], get_output(s))(a, b)
    assert np.array_equal(got, a[:, :, 1:7, 2:7] + b)                            # offset (size - min) // 2 per axis
    c = ConcatLayer((lb, la), axis=1, cropping=[None, None, 'center', 'center'])
    got = theano.function([av, bv], get_output(c))(a, b)
    assert np.array_equal(got, np.concatenate([b, a[:, :, 1:7, 2:7]], axis=1))  # incomings in the order given
    assert c.output_shape == (None, 4, None, None)
    pad = PadLayer(la, 3)
    got = theano.function([av], get_output(pad))(a)
    assert got.shape == (1, 2, 15, 16) and np.array_equal(got[:, :, 3:-3, 3:-3], a) and float(np.abs(got).sum() - np.abs(a).sum()) < 1e-3


def test_graph_traversal_order_is_lasagnes():
    """get_all_layers is a depth-first post-order over input_layer / input_layers (first incoming first); parameters follow it,
    W before b.  This order is the positional checkpoint format."""
    xv = T.tensor4('x')
    lin = InputLayer((None, 3, None, None), xv, name='in')
    a = Conv2DLayer(lin, 2, 3, pad='same', name='a')
    b = Conv2DLayer(a, 2, 3, pad='same', name='b')
    c = Conv2DLayer(lin, 2, 3, pad='same', name='c')
    s = ElemwiseSumLayer((c, b), name='s')
    assert [l.name for l in get_all_layers(s)] == ['in', 'c', 'a', 'b', 's']
    assert [p.name for p in get_all_params(s)] == ['c.W', 'c.b', 'a.W', 'a.b', 'b.W', 'b.b']
    with pytest.raises(ValueError):
        lasagne.layers.set_all_param_values(s, [np.zeros((2, 3, 3, 3), np.float32)] * 5)
    with pytest.raises(ValueError):
        lasagne.layers.set_all_param_values(s, [np.zeros((1,), np.float32)] * 6)


def test_symbolic_front_end_pieces_the_reference_uses():
    x = RNG.rand(2, 3, 4, 5).astype(np.float32)
    xv = T.tensor4('x')
    d = xv.dimshuffle((0, 2, 3, 1))
    sh = d.shape
    two_d = d.reshape((T.prod(sh[:3]), sh[3]))
    assert two_d.ndim == 2
    got = theano.function([xv], two_d)(x)
    assert np.array_equal(got, x.transpose(0, 2, 3, 1).reshape(-1, 3))
    cm = T.zeros((3, 3))
    pred = T.argmax(two_d, axis=1)
    for i in range(3):
        cm = T.set_subtensor(cm[i, i], T.sum(T.eq(pred, i) * T.eq(pred, i)))
    assert np.array_equal(theano.function([xv], cm)(x).diagonal(), np.bincount(got.argmax(1), minlength=3))
    with pytest.raises(TypeError):
        bool(pred < 1)                                                           # results of python comparisons are not truthy in Theano
    assert bool(pred.nonzero()[0])                                               # ... every other variable is
    sm = theano.function([xv], T.nnet.softmax(two_d))(x)
    e = np.exp(got - got.max(1, keepdims=True))
    assert np.abs(sm - e / e.sum(1, keepdims=True)).max() < 1e-6
    y = (xv * xv).sum()
    assert np.abs(theano.function([xv], theano.grad(y, xv))(x) - 2 * x).max() < 1e-6
    w = theano.shared(np.float32(2.0))
    f = theano.function([xv], (w * xv).sum(), updates={w: w + 1})
    assert abs(float(f(x)) - 2 * x.sum()) < 1e-3 and abs(float(f(x)) - 3 * x.sum()) < 1e-3 and float(w.get_value()) == 4.0


def test_python2_source_rewriting():
    src = '\n'.join([
        "print 'a', 1",
        "print('%d' % 3)",
        "print 'x {}'.format(",
        "    5)",
        "d = dict({'a': 1}.items())",
        "f = {'b': 2}",
        "e = dict(d.items() + f.items())",
        "q = 7 / 2",
        "r = 7.0 / 2",
        "n = [i for i in xrange(3)]",
    ])
    import ast
    tree = py2import._Div().visit(ast.parse(py2import.convert(src)))
    ast.fix_missing_locations(tree)
    ns = {'__py2div__': py2import.py2div, 'xrange': range}
    exec(compile(tree, '<py2>', 'exec'), ns)
    assert ns['q'] == 3 and ns['r'] == 3.5 and ns['e'] == {'a': 1, 'b': 2} and ns['n'] == [0, 1, 2]
    assert py2import.py2div(np.int64(7), 2) == 3 and py2import.py2div(7, 2.0) == 3.5

"""FC-DenseNet103 forward restated on torch CPU (test oracle; nothing here ships).

Follows models/FCDenseNet.py:15-146,196-219 (the network wiring) and, for the four layer helpers it
imports from the un-vendored, unpinned dependency `FC_DenseNet.layers` (SimJeg/FC-DenseNet,
models/FCDenseNet.py:12), the published recipe of that repository / arXiv 1611.09326:

  BN_ReLU_Conv(x, n, k=3)   = BatchNormLayer -> rectify -> Conv2DLayer(n, k, pad='same', linear,
                              HeUniform(gain='relu'), flip_filters=False) [-> Dropout: identity at inference]
  TransitionDown(x, n)      = BN_ReLU_Conv(x, n, k=1) -> Pool2DLayer(2, 'max')
  TransitionUp(skip, blk,k) = ConcatLayer(blk) -> Deconv2DLayer(k, 3, stride=2, crop='valid', linear)
                              -> ConcatLayer([deconv, skip], cropping=center)   (deconv output FIRST)
  SoftmaxLayer(x, C)        = Conv2DLayer(C, 1, linear) -> NHWC reshape -> softmax

Lasagne BatchNormLayer: (x - mean) * (gamma * inv_std) + beta, inv_std = 1/sqrt(var + 1e-4), biased
variance over axes (0,2,3).  The scripts evaluate with batch_norm_use_averages=False
(iterative_inference.py:187), i.e. BATCH statistics even at test time, so the stored running mean /
inv_std are never read; results depend on the batch composition.

PINNING: the network wiring, the hidden outputs and the 590-array positional checkpoint order are checked against the
reference's own models/FCDenseNet.py (Network / build_fcdensenet / restore), executed through oracle/refrun
(tests/golden/ref_densenet.npz, tests/test_oracle.py::test_oracle_vs_reference_run[ref_densenet], 3e-7).  What stays restated:
the four helpers of FC_DenseNet.layers above (the package is not in the tree; oracle/refrun/stubs/FC_DenseNet/layers.py
restates them a second time on the Lasagne stand-in).  Parameters are a flat list in creation order (= Lasagne's topological
order for this graph): conv W,b ; per BN_ReLU_Conv: beta, gamma, mean, inv_std, W, b ; deconv W,b.
"""
import torch
import torch.nn.functional as F

from . import lasagne_semantics as L

N_LAYERS_103 = [4, 5, 7, 10, 12, 15, 12, 10, 7, 5, 4]
BN_EPS = 1e-4


def densenet_param_shapes(nb_in_channels=3, n_classes=11, n_first=48, n_pool=5, growth=16,
                          n_layers=N_LAYERS_103):
    """[(name, kind, W shape)] in checkpoint order; kind in {'conv', 'bnconv', 'deconv'}.
    'conv' -> W, b; 'bnconv' -> beta, gamma, mean, inv_std, W, b; 'deconv' -> W (in,out,3,3), b."""
    out = [('first_conv', 'conv', (n_first, nb_in_channels, 3, 3))]
    n = n_first
    skips = []
    for i in range(n_pool):
        for j in range(n_layers[i]):
            out.append(('down%d_l%d' % (i, j), 'bnconv', (growth, n, 3, 3)))
            n += growth
        skips.append(n)
        out.append(('td%d' % i, 'bnconv', (n, n, 1, 1)))
    skips = skips[::-1]
    for j in range(n_layers[n_pool]):
        out.append(('bottleneck_l%d' % j, 'bnconv', (growth, n, 3, 3)))
        n += growth
    up_ch = growth * n_layers[n_pool]
    for i in range(n_pool):
        keep = growth * n_layers[n_pool + i]
        out.append(('tu%d' % i, 'deconv', (up_ch, keep, 3, 3)))
        n = keep + skips[i]
        for j in range(n_layers[n_pool + i + 1]):
            out.append(('up%d_l%d' % (i, j), 'bnconv', (growth, n, 3, 3)))
            n += growth
        up_ch = growth * n_layers[n_pool + i + 1]
    out.append(('softmax_conv', 'conv', (n_classes, n, 1, 1)))
    return out


def synthetic_densenet_params(nb_in_channels=3, n_classes=11, seed=2, logit_gain=1.0, **kw):
    """HeUniform conv / deconv W (lasagne: a = sqrt(6 / fan_in), fan_in = prod(shape[1:])), zero b,
    BN beta = 0, gamma = 1, mean = 0, inv_std = 1 (lasagne defaults); `logit_gain` scales the
    final 1x1 conv so y0 is peaky (the temperature lever of the FCN8 recipe)."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, kind, ws in densenet_param_shapes(nb_in_channels, n_classes, **kw):
        a = (6.0 / (ws[1] * ws[2] * ws[3])) ** 0.5
        W = (torch.rand(ws, generator=gen, dtype=torch.float32) * 2 - 1) * a
        if name == 'softmax_conv':
            W = W * logit_gain
        cin = ws[1] if kind != 'deconv' else ws[0]
        cout = ws[0] if kind != 'deconv' else ws[1]
        if kind == 'bnconv':
            params += [torch.zeros(cin), torch.ones(cin), torch.zeros(cin), torch.ones(cin)]
        params += [W, torch.zeros(cout)]
    return params


def _bn_relu_conv(x, beta, gamma, W, b):
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    inv_std = 1.0 / torch.sqrt(var + BN_EPS)
    xn = (x - mean) * (gamma.view(1, -1, 1, 1) * inv_std) + beta.view(1, -1, 1, 1)
    return L.conv2d(torch.relu(xn), W, b, pad='same', relu=False)


def densenet_forward(params, X, n_classes=11, layer=('pool4',), n_first=48, n_pool=5, growth=16,
                     n_layers=N_LAYERS_103):
    """models/FCDenseNet.py:61-146.  Returns hidden_outputs (the stacks after the named
    TransitionDowns, models/FCDenseNet.py:96-97) + [channel softmax (B, n_classes, H, W)]."""
    it = iter(params)

    def conv():
        return next(it), next(it)

    def bnconv():
        beta, gamma, _mean, _inv_std = next(it), next(it), next(it), next(it)
        W, b = next(it), next(it)
        return beta, gamma, W, b

    hidden_ints = [int(h[-1]) for h in layer if h.startswith('pool')]
    hidden = []
    W, b = conv()
    stack = L.conv2d(X, W, b, pad='same', relu=False)
    skips = []
    for i in range(n_pool):
        for j in range(n_layers[i]):
            stack = torch.cat([stack, _bn_relu_conv(stack, *bnconv())], dim=1)
        skips.append(stack)
        stack = L.maxpool2(_bn_relu_conv(stack, *bnconv()))
        if i + 1 in hidden_ints:
            hidden.append(stack)
    skips = skips[::-1]
    block = []
    for j in range(n_layers[n_pool]):
        l = _bn_relu_conv(stack, *bnconv())
        block.append(l)
        stack = torch.cat([stack, l], dim=1)
    for i in range(n_pool):
        W, b = conv()
        l = L.deconv2d(torch.cat(block, dim=1), W, b, stride=2)
        a, s = L.center_crop_pair(l, skips[i])
        stack = torch.cat([a, s], dim=1)
        block = []
        for j in range(n_layers[n_pool + i + 1]):
            l = _bn_relu_conv(stack, *bnconv())
            block.append(l)
            stack = torch.cat([stack, l], dim=1)
    W, b = conv()
    logits = L.conv2d(stack, W, b, pad='same', relu=False)
    assert next(it, None) is None, 'unused parameters'
    return hidden + [L.channel_softmax(logits)]

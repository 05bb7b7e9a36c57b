"""Synthetic CamVid-shaped inputs and random-init weights (the benchmark recipe).

There is no network for datasets or checkpoints, so bench.py and smoke runs use seeded synthetic
data of the reference's shapes: images U[0,1) (B,3,H,W) (`return_0_255=False`,
iterative_inference.py:117-118), one-hot float32 targets with a void channel (B,C+1,H,W)
(`target_var[:, :void]`, iterative_inference.py:125), and weights drawn with Lasagne's initialisers
in the positional checkpoint order of models/fcn8.py:177-180 / models/DAE_h.py:52-57.  The recipe
(He-uniform FCN8 with the `upsample` kernel scaled so y0 is peaky, GlorotUniform zero-bias DAE with
the last conv scaled so the iterated map is contractive) is part of the benchmark definition; the
test oracle carries an independent copy (oracle/weights.py) and tests/test_host_logic.py checks
that both produce identical tensors.
"""
import numpy as np
import torch

_VGG = [('conv1_1', 64), ('conv1_2', 64), ('conv2_1', 128), ('conv2_2', 128), ('conv3_1', 256), ('conv3_2', 256),
        ('conv3_3', 256), ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512), ('conv5_1', 512), ('conv5_2', 512),
        ('conv5_3', 512)]


def fcn8_shapes(nb_in_channels, n_classes, concat=None):
    """(name, W shape, b shape) of the 21 parameterised FCN8 layers (models/fcn8.py:33-110).  `concat` = (layer, nb_h): the
    FCN8-shaped DAE's conditioning channels in front of the named layer's consumer (models/fcn8_dae.py:46-115)."""
    out, cin = [], nb_in_channels
    if concat is not None and concat[0] == 'input':
        cin += concat[1]
    last_of_stage = {'conv1_2': 'pool1', 'conv2_2': 'pool2', 'conv3_3': 'pool3', 'conv4_3': 'pool4', 'conv5_3': 'pool5'}
    for name, cout in _VGG:
        out.append((name, (cout, cin, 3, 3), (cout,)))
        cin = cout
        if concat is not None and last_of_stage.get(name) == concat[0]:
            cin += concat[1]
    c = n_classes
    out += [('fc6', (4096, 512, 7, 7), (4096,)), ('fc7', (4096, 4096, 1, 1), (4096,)),
            ('score_fr', (c, 4096, 1, 1), (c,)), ('score2', (c, c, 4, 4), (c,)),
            ('score_pool4', (c, 512, 1, 1), (c,)), ('score4', (c, c, 4, 4), (c,)),
            ('score_pool3', (c, 256, 1, 1), (c,)), ('upsample', (c, c, 16, 16), (c,))]
    return out


def dae_shapes(n_classes, nb_features_to_concat, n_filters=64, concat_h=('pool4',), additional_pool=2, unpool_type='trackind'):
    """(name, W shape, b shape) of DAE_h's convs: conv1_1..convP_1 (models/fcn_down.py:96-104) then
    up_convP..up_conv1 (models/fcn_up.py:29-34,84-86)."""
    last = concat_h[-1]
    n_pool = int(last[-1]) if 'pool' in last else 0
    total = n_pool + additional_pool
    out, widths, cin = [], [], n_classes + (nb_features_to_concat if last == 'input' else 0)
    f = n_filters
    for p in range(total):
        if p < 6:
            f = n_filters * 2 ** p
        out.append(('conv%d_1' % (p + 1), (f, cin, 3, 3), (f,)))
        widths.append(f)
        cin = f + (nb_features_to_concat if (p + 1 == n_pool and n_pool > 0) else 0)
    up_in = widths[-1]
    for p in range(total, 0, -1):
        n_cl = n_classes if p == 1 else widths[p - 2]
        if unpool_type == 'standard':      # Deconv2DLayer(n_cl, 4, stride=2): W (in, out, 4, 4), models/fcn_up.py:41-45
            out.append(('up%d' % p, (up_in, n_cl, 4, 4), (n_cl,)))
        else:
            out.append(('up_conv%d' % p, (n_cl, up_in, 3, 3), (n_cl,)))
        up_in = n_cl
    return out


def _uniform(shape, a, gen):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * a


def synthetic_fcn8_params(nb_in_channels, n_classes, seed=0, logit_gain=1.0, concat=None):
    """lasagne.init.HeUniform (a = sqrt(6 / fan_in)) W, zero b; `upsample` W x logit_gain."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, ws, bs in fcn8_shapes(nb_in_channels, n_classes, concat):
        W = _uniform(ws, np.sqrt(6.0 / int(np.prod(ws[1:]))), gen)
        params += [W * logit_gain if name == 'upsample' else W, torch.zeros(bs)]
    return params


def synthetic_dae_params(n_classes, nb_features_to_concat, seed=1, n_filters=64, concat_h=('pool4',),
                         additional_pool=2, out_gain=1.0, unpool_type='trackind'):
    """lasagne.init.GlorotUniform (a = sqrt(6 / ((n_out + n_in) * kh * kw))) W, zero b; `up_conv1` W x out_gain."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, ws, bs in dae_shapes(n_classes, nb_features_to_concat, n_filters, concat_h, additional_pool, unpool_type):
        W = _uniform(ws, np.sqrt(6.0 / ((ws[0] + ws[1]) * int(np.prod(ws[2:])))), gen)
        params += [W * out_gain if name in ('up_conv1', 'up1') else W, torch.zeros(bs)]
    return params


def synthetic_contextmod_params(n_classes, nb_features_to_concat=3, seed=3, out_gain=4.0):
    """Context module (models/contextmod_dae.py:72-103): conv1 (C, nb_h + C, 3, 3), dilconv1..6 (C, C, 3, 3) and dilconv7
    (C, C, 1, 1) in DilatedConv2DLayer's (in, out) order.  The reference starts the dilated convs at the identity
    (IdentityInit, :61-71); the stand-in for a trained module is identity centre taps + half-scale Glorot-uniform noise
    on every weight, biases U(-0.05, 0.05), the output conv x out_gain."""
    gen = torch.Generator().manual_seed(seed)
    C_ = n_classes
    shapes = [('conv1', (C_, nb_features_to_concat + C_, 3, 3))] + [('dilconv%d' % (i + 1), (C_, C_, 3, 3)) for i in range(6)] + \
             [('dilconv7', (C_, C_, 1, 1))]
    params = []
    for name, ws in shapes:
        W = _uniform(ws, np.sqrt(6.0 / ((ws[0] + ws[1]) * int(np.prod(ws[2:])))), gen) * 0.5
        if name != 'conv1':
            k = ws[2] // 2
            for i in range(ws[0]):
                W[i, i, k, k] += 1.0
        if name == 'dilconv7':
            W = W * out_gain
        params += [W, (torch.rand((C_,), generator=gen) - 0.5) * 0.1]
    return params


N_LAYERS_103 = [4, 5, 7, 10, 12, 15, 12, 10, 7, 5, 4]


def densenet_shapes(nb_in_channels=3, n_classes=11, n_first=48, n_pool=5, growth=16, n_layers=N_LAYERS_103):
    """(name, kind, W shape) of FC-DenseNet103's 103 parameterised layers in creation order
    (models/FCDenseNet.py:61-141): kind 'conv' -> W, b; 'bnconv' -> beta, gamma, mean, inv_std, W, b;
    'deconv' -> W (in, out, 3, 3), b."""
    out = [('first_conv', 'conv', (n_first, nb_in_channels, 3, 3))]
    n, skips = n_first, []
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


def synthetic_densenet_params(nb_in_channels=3, n_classes=11, seed=2, logit_gain=1.0):
    """HeUniform conv / deconv W, zero b, BatchNorm beta 0 / gamma 1 (lasagne defaults); the final 1x1 conv x logit_gain."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, kind, ws in densenet_shapes(nb_in_channels, n_classes):
        W = _uniform(ws, (6.0 / (ws[1] * ws[2] * ws[3])) ** 0.5, gen)
        if name == 'softmax_conv':
            W = W * logit_gain
        cin, cout = (ws[1], ws[0]) if kind != 'deconv' else (ws[0], ws[1])
        if kind == 'bnconv':
            params += [torch.zeros(cin), torch.ones(cin), torch.zeros(cin), torch.ones(cin)]
        params += [W, torch.zeros(cout)]
    return params


def synthetic_batch(B, H, W, n_classes=11, seed=0):
    """X (B,3,H,W) U[0,1); one-hot float32 targets (B, n_classes+1, H, W), label n_classes = void; labels."""
    gen = torch.Generator().manual_seed(seed)
    X = torch.rand((B, 3, H, W), generator=gen, dtype=torch.float32)
    lab = torch.randint(0, n_classes + 1, (B, H, W), generator=gen)
    L = torch.nn.functional.one_hot(lab, n_classes + 1).permute(0, 3, 1, 2).float()
    return X, L.contiguous(), lab

"""FCN8 and DAE_h forward passes restated on torch CPU (test oracle).

Parameters are flat lists in Lasagne `get_all_param_values` order (the order of
the reference's positional `.npz` checkpoints, models/DAE_h.py:52-57,
models/fcn8.py:177-180).
"""
import torch

from . import lasagne_semantics as L

# ---------------------------------------------------------------------------
# FCN8 (models/fcn8.py:16-200)
# ---------------------------------------------------------------------------

VGG_CFG = [  # (name, out_channels) grouped per pooling stage; models/fcn8.py:33-72
    [('conv1_1', 64), ('conv1_2', 64)],
    [('conv2_1', 128), ('conv2_2', 128)],
    [('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256)],
    [('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512)],
    [('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512)],
]


def fcn8_param_shapes(nb_in_channels, n_classes, concat=None):
    """[(name, W shape, b shape)] in checkpoint order: 13 VGG convs, fc6, fc7,
    score_fr, score2, score_pool4, score4, score_pool3, upsample (42 arrays).
    Conv W is (out,in,kh,kw); Deconv W is (in,out,kh,kw).
    `concat` = (layer name, nb_h): the FCN8-shaped DAE (models/fcn8_dae.py:46-48,60-115) concatenates nb_h conditioning
    channels BEFORE the named layer's output ('input', 'pool1'..'pool4'), widening the next conv's input."""
    shapes = []
    cin = nb_in_channels
    if concat is not None and concat[0] == 'input':
        cin += concat[1]
    for si, stage in enumerate(VGG_CFG):
        for name, cout in stage:
            shapes.append((name, (cout, cin, 3, 3), (cout,)))
            cin = cout
        if concat is not None and concat[0] == 'pool%d' % (si + 1):
            cin += concat[1]
    shapes.append(('fc6', (4096, 512, 7, 7), (4096,)))
    shapes.append(('fc7', (4096, 4096, 1, 1), (4096,)))
    shapes.append(('score_fr', (n_classes, 4096, 1, 1), (n_classes,)))
    shapes.append(('score2', (n_classes, n_classes, 4, 4), (n_classes,)))
    shapes.append(('score_pool4', (n_classes, 512, 1, 1), (n_classes,)))
    shapes.append(('score4', (n_classes, n_classes, 4, 4), (n_classes,)))
    shapes.append(('score_pool3', (n_classes, 256, 1, 1), (n_classes,)))
    shapes.append(('upsample', (n_classes, n_classes, 16, 16), (n_classes,)))
    return shapes


def fcn8_forward(params, X, n_classes, layer=('pool4', 'probs_dimshuffle'),
                 temperature=1.0, concat=None):
    """models/fcn8.py:30-130,187-200.  Returns [net[el] for el in layer].
    `temperature` divides upsample.W and .b (models/fcn8.py:193-198).
    Dropout layers are identity (deterministic=True, iterative_inference.py:187).
    NOTE score_fr / score_pool4 / score_pool3 keep Lasagne's default rectify."""
    names = [s[0] for s in fcn8_param_shapes(X.shape[1], n_classes)]
    P = {n: (params[2 * i], params[2 * i + 1]) for i, n in enumerate(names)}
    net = {}
    x = X
    if concat is not None and concat[0] == 'input':       # (h, layer): ConcatLayer((h, net[layer])), models/model_helpers.py:91-93
        x = torch.cat([concat[1], x], dim=1)
    for si, stage in enumerate(VGG_CFG):
        for ci, (name, _) in enumerate(stage):
            pad = 100 if name == 'conv1_1' else 'same'
            x = L.conv2d(x, *P[name], pad=pad, relu=True)
            net[name] = x
        x = L.maxpool2(x)
        net['pool%d' % (si + 1)] = x               # score_pool3 / score_pool4 read the un-concatenated pool
        if concat is not None and concat[0] == 'pool%d' % (si + 1):
            x = torch.cat([concat[1], x], dim=1)
    x = L.conv2d(x, *P['fc6'], pad='valid', relu=True)
    net['fc6'] = x
    x = L.conv2d(x, *P['fc7'], pad='valid', relu=True)
    net['fc7'] = x
    x = L.conv2d(x, *P['score_fr'], pad='valid', relu=True)
    net['score_fr'] = x
    s2 = L.deconv2d(x, *P['score2'], stride=2)
    net['score2'] = s2
    sp4 = L.conv2d(net['pool4'], *P['score_pool4'], pad='same', relu=True)
    a, b = L.center_crop_pair(s2, sp4)
    net['score_fused'] = a + b
    s4 = L.deconv2d(net['score_fused'], *P['score4'], stride=2)
    net['score4'] = s4
    sp3 = L.conv2d(net['pool3'], *P['score_pool3'], pad='valid', relu=True)
    a, b = L.center_crop_pair(s4, sp3)
    net['score_final'] = a + b
    Wu, bu = P['upsample']
    up = L.deconv2d(net['score_final'], Wu / temperature, bu / temperature, stride=8)
    net['upsample'] = up
    net['score'] = L.center_crop_to(up, X.shape[2], X.shape[3])
    net['probs_dimshuffle'] = L.channel_softmax(net['score'])
    return [net[el] for el in layer]


# ---------------------------------------------------------------------------
# DAE_h (models/DAE_h.py:12-63, models/fcn_down.py:9-138, models/fcn_up.py:11-172)
# kind='standard', unpool_type='trackind', conv_before_pool=1, skip=True, bn=0,
# dropout=0, noise=0 (parity is pinned at noise=0: layers/mylayers.py:91-93
# rebuilds the mask sub-graph non-deterministically otherwise).
# ---------------------------------------------------------------------------

def dae_levels(concat_h=('pool4',), additional_pool=2):
    """n_pool from the last concat name + additional_pool (models/DAE_h.py:36-39)."""
    last = concat_h[-1]
    n_pool = int(last[-1]) if 'pool' in last else 0
    return n_pool, n_pool + additional_pool


def dae_param_shapes(n_classes, nb_features_to_concat, n_filters=64,
                     concat_h=('pool4',), additional_pool=2, unpool_type='trackind', conv_before_pool=1):
    """[(name, W shape, b shape)] in checkpoint order: conv1_1..convP_1 then
    up_convP..up_conv1.  Filter counts: n_filters*2**p for p<6
    (models/fcn_down.py:96-99); up_conv_p outputs the channel count of
    pool_{p-1}'s input, or n_classes for p==1 (models/fcn_up.py:29-34)."""
    n_pool, total = dae_levels(concat_h, additional_pool)
    shapes = []
    cin = n_classes
    if concat_h[-1] == 'input':
        cin += nb_features_to_concat
    conv_out = []
    filters = n_filters
    for p in range(total):
        if p < 6:
            filters = n_filters * (2 ** p)
        for i in range(1, conv_before_pool + 1):       # models/fcn_down.py:83-104
            shapes.append(('conv%d_%d' % (p + 1, i), (filters, cin, 3, 3), (filters,)))
            cin = filters
        conv_out.append(filters)
        if p + 1 == n_pool and n_pool > 0:  # concat h after pool{n_pool}
            cin += nb_features_to_concat
    up_in = conv_out[-1]
    for p in range(total, 0, -1):
        # pool_{p-1}.input_shape[1]: the (un-concatenated) conv output of level p-1
        n_cl = n_classes if p == 1 else conv_out[p - 2]
        if unpool_type == 'standard':   # Deconv2DLayer(n_cl, 4, stride=2): W (in, out, 4, 4)  (models/fcn_up.py:41-45)
            shapes.append(('up%d' % p, (up_in, n_cl, 4, 4), (n_cl,)))
        else:
            shapes.append(('up_conv%d' % p, (n_cl, up_in, 3, 3), (n_cl,)))
        up_in = n_cl
    return shapes


def batchnorm_deterministic(x, beta, gamma, mean, inv_std):
    """lasagne BatchNormLayer under deterministic=True: (x - mean) * (gamma * inv_std) + beta on the STORED averages
    (parameter order beta, gamma, mean, inv_std).  Reference: models/fcn_down.py:113-115, models/fcn_up.py:91-93 with
    get_output(dae, deterministic=True), iterative_inference.py:189-190."""
    v = lambda a: a.view(1, -1, 1, 1)          # noqa: E731
    return (x - v(mean)) * (v(gamma) * v(inv_std)) + v(beta)


def batchnorm_batch_stats(x, beta, gamma, epsilon=1e-4):
    """lasagne BatchNormLayer with deterministic=False: normalises with the mean and the biased variance of the current
    batch over (batch, rows, cols), inv_std = 1 / sqrt(var + 1e-4); the stored averages are not read (and not updated: the
    reference collects no updates at inference)."""
    v = lambda a: a.view(1, -1, 1, 1)          # noqa: E731
    mean = x.mean(dim=(0, 2, 3))
    inv_std = 1.0 / torch.sqrt(x.var(dim=(0, 2, 3), unbiased=False) + epsilon)
    return (x - v(mean)) * (v(gamma) * v(inv_std)) + v(beta)


def dae_forward(params, y, h, padding, concat_h=('pool4',), additional_pool=2,
                return_logits=False, unpool_type='trackind', bn=False, mask_source_y=None, skip=True,
                conv_before_pool=1):
    """One application DAE(y, h) -> probabilities, same size as y.

    Down (models/fcn_down.py:77-136): conv3x3 ReLU (pad=`padding` on the first
    conv when concatenating at a pool layer and padding>0, else 'same'),
    maxpool2; h is concatenated BEFORE the DAE's own pool{n_pool} channels
    (models/model_helpers.py:93-94).
    Up (models/fcn_up.py:65-113): DePool2D with the tie mask of level p's
    pre-pool map, conv3x3 'same' linear, skip-sum with the un-concatenated
    pool_{p-1} (p>1) or centre crop to the input size (p==1); channel softmax.
    unpool_type='inverse' (models/fcn_up.py:76-79): lasagne InverseLayer(prev, pool_p) = the gradient of
    the max-pool w.r.t. its input with `prev` as the upstream gradient; with Theano's CPU MaxPoolGrad
    (every tied maximum receives it) that is exactly DePool2D's repeat * tie-mask, so it shares the code.
    mask_source_y: the reference builds DePool2D's mask sub-graph with get_output(...) WITHOUT deterministic=True
    (layers/mylayers.py:91-93): when the DAE has noise > 0 its tie masks come from a SEPARATE pass of the
    contracting path on y + N(0, noise^2), even at inference.  Pass that noised input here to restate it
    (None: the deterministic graph, masks from the same pass; a LIST of `total` noised inputs: the reference's actual graph,
    an independent noise draw per DePool2D, level p's mask from a pass on entry p - 1).
    unpool_type='standard' (models/fcn_up.py:37-63): up_p = Deconv2DLayer(prev, n_cl, 4, stride=2,
    crop='valid', linear) and NO convolution; skip-sum / crop as above (centre crop of the larger map).
    """
    n_pool, total = dae_levels(concat_h, additional_pool)
    if bn:     # bn=1: a BatchNormLayer (4 arrays) behind every conv of both paths, except after 'standard' deconvs
        k_up = 2 if unpool_type == 'standard' else 6
        Wd = [tuple(params[6 * i:6 * i + 2]) for i in range(total)]
        BNd = [tuple(params[6 * i + 2:6 * i + 6]) for i in range(total)]
        Wu = [tuple(params[6 * total + k_up * i:6 * total + k_up * i + 2]) for i in range(total)]
        BNu = [tuple(params[6 * total + k_up * i + 2:6 * total + k_up * i + 6]) if k_up == 6 else None for i in range(total)]
    elif conv_before_pool > 1:     # k convs per level (models/fcn_down.py:83-104): the first as below, then 'same' convs
        k = conv_before_pool
        Wd = [(params[2 * k * i], params[2 * k * i + 1]) for i in range(total)]
        Wd_extra = [[(params[2 * k * i + 2 * j], params[2 * k * i + 2 * j + 1]) for j in range(1, k)] for i in range(total)]
        Wu = [(params[2 * (k * total + i)], params[2 * (k * total + i) + 1]) for i in range(total)]
        BNd = BNu = [None] * total
    else:
        Wd = [(params[2 * i], params[2 * i + 1]) for i in range(total)]
        Wu = [(params[2 * (total + i)], params[2 * (total + i) + 1]) for i in range(total)]
        BNd = BNu = [None] * total
    if conv_before_pool <= 1:
        Wd_extra = [[] for _ in range(total)]
    x = y
    if concat_h[-1] == 'input':
        x = torch.cat([h, x], dim=1)
    pre, pools = [], []
    for p in range(total):
        first_pad = (p == 0 and len(concat_h) == 1 and concat_h[-1] != 'input'
                     and padding > 0)
        x = L.conv2d(x, *Wd[p], pad=padding if first_pad else 'same', relu=True)
        for We in Wd_extra[p]:
            x = L.conv2d(x, *We, pad='same', relu=True)
        if BNd[p] is not None:
            x = batchnorm_deterministic(x, *BNd[p])
        pre.append(x)
        x = L.maxpool2(x)
        pools.append(x)
        if p + 1 == n_pool and n_pool > 0:
            x = torch.cat([h, x], dim=1)
    if mask_source_y is not None or (bn and unpool_type == 'trackind'):
        # DePool2D's own sub-graph, built by get_output(...) WITHOUT deterministic=True (layers/mylayers.py:91-93): same
        # weights, the noised input when noise > 0, and -- deterministic being False there -- every BatchNormLayer on the
        # statistics of ITS OWN BATCH (lasagne: batch_norm_use_averages defaults to `deterministic`), not on the stored
        # averages.  Found by running the reference (tests/golden/ref_bn.npz); 'inverse' (InverseLayer) receives the
        # deterministic expressions and is not affected.
        def mask_pass(xm, upto):
            out = []
            if concat_h[-1] == 'input':
                xm = torch.cat([h, xm], dim=1)
            for p in range(upto):
                first_pad = (p == 0 and len(concat_h) == 1 and concat_h[-1] != 'input' and padding > 0)
                xm = L.conv2d(xm, *Wd[p], pad=padding if first_pad else 'same', relu=True)
                for We in Wd_extra[p]:
                    xm = L.conv2d(xm, *We, pad='same', relu=True)
                if BNd[p] is not None:
                    xm = batchnorm_batch_stats(xm, BNd[p][0], BNd[p][1])
                out.append(xm)
                xm = L.maxpool2(xm)
                if p + 1 == n_pool and n_pool > 0:
                    xm = torch.cat([h, xm], dim=1)
            return out
        if isinstance(mask_source_y, (list, tuple)):
            # one noised input PER DePool2D: every DePool2D calls get_output(...) itself (layers/mylayers.py:91-93), and every
            # symbolic call of GaussianNoiseLayer is a new, independent random stream in Theano -- level p's mask comes from its
            # own pass on y + N_p (observed by executing the reference with logged draws: tests/golden/ref_noise.npz)
            assert len(mask_source_y) == total
            pre = [mask_pass(mask_source_y[p], p + 1)[p] for p in range(total)]
        else:
            pre = mask_pass(y if mask_source_y is None else mask_source_y, total)
    u = pools[-1]
    for i, p in enumerate(range(total, 0, -1)):
        if unpool_type == 'standard':
            u = L.deconv2d(u, *Wu[i], stride=2)
        elif unpool_type in ('trackind', 'inverse'):
            u = L.depool2d(u, pre[p - 1])
            u = L.conv2d(u, *Wu[i], pad='same', relu=False)
            if BNu[i] is not None:
                u = batchnorm_deterministic(u, *BNu[i])
        else:
            raise ValueError('Unkown unpool type')
        if p > 1:
            a, b = L.center_crop_pair(u, pools[p - 2])
            u = a + b if skip else a          # skip=False: CroppingLayer keeps the (cropped) up-conv only, models/fcn_up.py:103-113
        else:
            u = L.center_crop_to(u, y.shape[2], y.shape[3])
    if return_logits:
        return u
    return L.channel_softmax(u)


# ---------------------------------------------------------------------------
# Context-module DAE (models/contextmod_dae.py:19-138), kind='contextmod'
# ---------------------------------------------------------------------------

CONTEXTMOD_DILATIONS = (1, 2, 4, 8, 16, 1)


def contextmod_param_shapes(n_classes, nb_features_to_concat=3):
    """[(name, W shape, b shape)] in checkpoint order: conv1 is a Conv2DLayer (out, in, 3, 3) on [h | y]; dilconv1..7 are
    DilatedConv2DLayers, W (in, out, k, k), the last one 1x1 (models/contextmod_dae.py:72-103)."""
    C = n_classes
    shapes = [('conv1', (C, nb_features_to_concat + C, 3, 3), (C,))]
    shapes += [('dilconv%d' % (i + 1), (C, C, 3, 3), (C,)) for i in range(6)]
    shapes.append(('dilconv7', (C, C, 1, 1), (C,)))
    return shapes


def contextmod_forward(params, y, h, return_logits=False):
    """One application of the context module -> probabilities, same size as y.  deterministic=True: the
    GaussianNoiseLayer on y is the identity (models/contextmod_dae.py:50-57); h (the image) is concatenated BEFORE y
    (models/model_helpers.py:91-93); conv1 'same' rectify, PadLayer(32), six dilated 3x3 rectify convs, a 1x1 linear one,
    channel softmax (out_nonlin=softmax, iterative_inference.py:176)."""
    x = L.conv2d(torch.cat([h, y], dim=1), params[0], params[1], 'same', relu=True)
    x = L.pad_layer(x, 32)
    for i, d in enumerate(CONTEXTMOD_DILATIONS):
        x = L.dilated_conv2d(x, params[2 + 2 * i], params[3 + 2 * i], d, relu=True)
    x = L.dilated_conv2d(x, params[14], params[15], 1, relu=False)
    assert x.shape[2:] == y.shape[2:]
    return x if return_logits else L.channel_softmax(x)


def fcn8_dae_forward(params, y, h, n_classes, concat_h=('pool4',)):
    """kind='fcn8' (models/fcn8_dae.py:19-271): the FCN8 graph with y as its input (nb_in_channels = n_classes,
    iterative_inference.py:166-170) and h concatenated before y at 'input' or before pool_N's output at 'poolN' -- layer for
    layer models/fcn8.py otherwise (noise and dropout are the identity under deterministic=True).  -> probabilities."""
    assert len(concat_h) == 1
    return fcn8_forward(params, y, n_classes, layer=('probs_dimshuffle',), concat=(concat_h[0], h))[0]

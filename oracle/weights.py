"""Lasagne-style initialisers, positional .npz checkpoints and the synthetic
benchmark recipe (test oracle; also the single source of synthetic weights so
the CUDA path and the oracle see identical numbers).

Checkpoint format: np.savez(path, *get_all_param_values(net)) -> arr_0..arr_k
(train_dae.py:436-445; loaded at models/DAE_h.py:52-57, models/fcn8.py:177-180).
"""
import numpy as np
import torch

from .nets import contextmod_param_shapes, dae_param_shapes, fcn8_param_shapes


def glorot_uniform(shape, gen):
    """lasagne.init.GlorotUniform: a = sqrt(6 / ((n1 + n2) * receptive_field)),
    n1, n2 = shape[:2]."""
    rf = int(np.prod(shape[2:]))
    a = np.sqrt(6.0 / ((shape[0] + shape[1]) * rf))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * a


def he_uniform(shape, gen):
    """lasagne.init.HeUniform(gain='relu'): a = sqrt(6 / fan_in),
    fan_in = prod(shape[1:])."""
    a = np.sqrt(6.0 / int(np.prod(shape[1:])))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * a


def synthetic_fcn8_params(nb_in_channels, n_classes, seed=0, logit_gain=1.0, concat=None):
    """He-uniform W, zero b.  With random weights the FCN8 logits are tiny, so
    `logit_gain` rescales the final `upsample` kernel to give peaky y0 (the
    same lever as temperature<1, models/fcn8.py:193-198)."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, ws, bs in fcn8_param_shapes(nb_in_channels, n_classes, concat):
        W = he_uniform(ws, gen)
        if name == 'upsample':
            W = W * logit_gain
        params += [W, torch.zeros(bs)]
    return params


def synthetic_dae_params(n_classes, nb_features_to_concat, seed=1, n_filters=64,
                         concat_h=('pool4',), additional_pool=2, out_gain=1.0, unpool_type='trackind', conv_before_pool=1):
    """Lasagne defaults: GlorotUniform W, zero b (the benign
    regime).  `out_gain` rescales the last conv (up_conv1): with random weights
    the iterated map amplifies pool-mask flips, and out_gain < 1 makes it
    contractive so that free-running parity is well defined
    (oracle/validate_recipe.py measures the fp32-vs-fp64 drift)."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, ws, bs in dae_param_shapes(n_classes, nb_features_to_concat, n_filters,
                                         concat_h, additional_pool, unpool_type, conv_before_pool):
        W = glorot_uniform(ws, gen)
        if name in ('up_conv1', 'up1'):
            W = W * out_gain
        params += [W, torch.zeros(bs)]
    return params


def synthetic_contextmod_params(n_classes, nb_features_to_concat=3, seed=3, out_gain=4.0):
    """The reference initialises the dilated convs to the identity (IdentityInit, models/contextmod_dae.py:61-71) and
    trains from there; a trained module is identity + a learned perturbation.  Synthetic stand-in: identity centre taps
    plus Glorot-uniform noise on every weight (so each tap, channel pair and dilation contributes), small random biases,
    `out_gain` on the 1x1 output conv for peaky probabilities."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    for name, ws, bs in contextmod_param_shapes(n_classes, nb_features_to_concat):
        W = glorot_uniform(ws, gen) * 0.5
        if name != 'conv1':
            k = ws[2] // 2
            for i in range(ws[0]):
                W[i, i, k, k] += 1.0
        if name == 'dilconv7':
            W = W * out_gain
        params += [W, (torch.rand(bs, generator=gen) - 0.5) * 0.1]
    return params


def save_npz(path, params):
    np.savez(path, *[np.asarray(p, dtype=np.float32) for p in params])


def load_npz(path):
    with np.load(path) as f:
        return [torch.from_numpy(f['arr_%d' % i]) for i in range(len(f.files))]


def synthetic_batch(B, H, W, n_classes=11, seed=0):
    """CamVid-shaped synthetic batch: X ~ U[0,1) (B,3,H,W);
    labels randint(0, n_classes+1) with n_classes = void, one-hot float32
    (B, n_classes+1, H, W)."""
    gen = torch.Generator().manual_seed(seed)
    X = torch.rand((B, 3, H, W), generator=gen, dtype=torch.float32)
    lab = torch.randint(0, n_classes + 1, (B, H, W), generator=gen)
    L = torch.nn.functional.one_hot(lab, n_classes + 1).permute(0, 3, 1, 2).float()
    return X, L.contiguous(), lab


def with_batchnorm(pd, n_levels, unpool_type, seed=21):
    """Insert BatchNormLayer parameters (beta, gamma, mean, inv_std -- lasagne's order) behind every conv of a DAE_h
    checkpoint, as bn=1 saves them (models/fcn_down.py:113-115, models/fcn_up.py:91-93; none behind the 'standard'
    deconvolutions); a few negative gammas so that nothing relies on a sign."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for i in range(2 * n_levels):
        W, b = pd[2 * i], pd[2 * i + 1]
        out += [W, b]
        if i >= n_levels and unpool_type == 'standard':
            continue
        c = b.shape[0]
        gamma = 0.5 + torch.rand(c, generator=gen)
        gamma[::7] *= -1.0
        out += [0.1 * torch.randn(c, generator=gen), gamma, 0.05 * torch.randn(c, generator=gen), 0.5 + 1.5 * torch.rand(c, generator=gen)]
    return out

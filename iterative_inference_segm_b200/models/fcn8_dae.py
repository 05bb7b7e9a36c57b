"""FCN8-shaped DAE (kind='fcn8'), drop-in for models/fcn8_dae.py:19-271.

The DAE is the FCN8 graph of models/fcn8.py with y as its input (nb_in_channels = n_classes,
iterative_inference.py:166-170) and the conditioning tensor h concatenated in front of y ('input') or of pool_N's output
('pool1'..'pool4'); GaussianNoiseLayer and the dropout layers are the identity under deterministic=True.  It runs on
FCN8Net's kernels (tcgen05 convs with fused pools, the <=16-channel deconvs); the iteration-invariant h half of the
widened conv is hoisted out of the loop (FCN8Net, `concat=`).

NB the reference declares the concatenated InputLayer with the channel count of the layer it joins
(models/fcn8_dae.py:46-48,60-62: `net['input'].output_shape[1]`, `net['pool1'].output_shape[1]`, ...), so its graphs only
build when h has as many channels as that layer: a pool_N of the segmentation FCN8 for 'poolN' (the same VGG stage), an
n_classes-channel tensor for 'input'.  `nb_features_to_concat` is explicit here.
"""
import os

import torch

from .. import _kernels as K
from .._packing import load_npz_params
from .fcn8 import FCN8Net, LayerHandle, _POOL_CHANNELS


class FCN8DaeNet(object):
    fusable_update = False            # the loop runs its stand-alone softmax / update kernel on the logits
    mask_noise = 0.0

    def __init__(self, n_classes, nb_h, params, concat_h=('pool4',), precision='bf16', device='cuda'):
        assert len(concat_h) == 1, 'one conditioning tensor'
        self.fcn = FCN8Net(n_classes, n_classes, params, device=device, precision=precision, concat=(concat_h[0], nb_h))
        self.concat_at = concat_h[0]
        self.n_classes, self.nb_h = n_classes, nb_h
        self.split, self.cm = self.fcn.split, self.fcn.cm
        self.h_pad, self.y_cpad = self.fcn.h_pad, K.pad_channels(n_classes, narrow=True)
        self.device = self.fcn.device

    def h_spatial(self, H, W):
        if self.concat_at == 'input':
            return H, W
        h, w = H + 198, W + 198                       # conv1_1 pad=100
        for _ in range(int(self.concat_at[-1])):
            h, w = h // 2, w // 2
        return h, w

    def logits(self, h_bf16, y_bf16, full_down=True, update=None, y_f32=None, noise=None):
        """h_bf16 / y_bf16: NHWC bf16 (pairs).  `full_down`: h is new -- recompute the hoisted W_h * h term.
        Returns the fp32 NHWC16 logits (B, H, W, 16)."""
        assert update is None
        return self.fcn.forward(None, want=('logits',), x_packed=y_bf16, h=h_bf16 if full_down else None)['logits']


def buildFCN8_DAE(input_concat_h_vars, input_mask_var, n_classes, nb_in_channels=3,
                  path_weights='/Tmp/romerosa/itinf/models/', model_name='fcn8_model.npz', trainable=False,
                  load_weights=False, pretrained=False, freeze=False, pretrained_path='', pascal=False,
                  return_layer='probs_dimshuffle', concat_h=['input'], noise=0.1, dropout=0.5, params=None,
                  precision='bf16', nb_features_to_concat=None):
    """Same arguments as the reference builder (models/fcn8_dae.py:19-25); returns the handle of 'probs_dimshuffle'.
    Inference only: weights come from `params` or the positional checkpoint (load_weights); the `pretrained` / `pascal`
    initialisations belong to training.  `nb_in_channels` is the channel count of y (the reference passes n_classes);
    `nb_features_to_concat` defaults to what the reference's graph implies (see the module docstring)."""
    if return_layer != 'probs_dimshuffle':
        raise NotImplementedError('B200 FCN8 DAE returns probs_dimshuffle')
    concat_h = list(concat_h)
    assert all(el in ['pool1', 'pool2', 'pool3', 'pool4', 'input'] for el in concat_h)           # models/fcn8_dae.py:34-35
    if len(concat_h) != 1:
        raise NotImplementedError('B200 FCN8 DAE concatenates one conditioning tensor')
    if nb_in_channels != n_classes:
        raise NotImplementedError('the DAE\'s input is y: nb_in_channels must equal n_classes')
    if nb_features_to_concat is None:
        nb_features_to_concat = nb_in_channels if concat_h[0] == 'input' else _POOL_CHANNELS[concat_h[0]]
    if params is None:
        if not load_weights:
            raise ValueError('buildFCN8_DAE needs weights: pass params= or load_weights=True with path_weights')
        params = load_npz_params(os.path.join(path_weights, model_name))
    net = FCN8DaeNet(n_classes, nb_features_to_concat, params, tuple(concat_h), precision=precision)
    return LayerHandle(net, 'probs_dimshuffle', n_classes)

"""Context-module DAE (kind='contextmod'), drop-in for models/contextmod_dae.py:19-138.

    [h | y] -> conv1 3x3 'same' rectify -> PadLayer(32) -> DilatedConv2DLayer 3x3, dilation 1, 2, 4, 8, 16, 1
            ('valid', rectify) -> DilatedConv2DLayer 1x1 (linear) -> channel softmax (out_nonlin)

h is the image (concat_h=['input'], the only value the reference accepts: models/contextmod_dae.py:42); the context
module keeps the image resolution.  Every layer has n_classes channels, so this is fp32 CUDA-core work (csrc/contextmod.cu),
exact float32 like the reference: fp32 activations (the master y and the image are read as planar NCHW, the module's own
tensors are channels-last with 12 channels), the weights in the kernels' parameter blocks.  Per application:
conv1 (+ the hoisted, iteration-invariant W_h * h term, computed once per batch) writes into the interior of the
zero-bordered PadLayer buffer, five dilated convs ping-pong between two buffers, and the sixth carries the 1x1 conv and
writes the fp32 NHWC16 logits the loop's softmax / update kernel consumes.  Frozen images (active == 0) are skipped.

Checkpoint order (lasagne.layers.get_all_param_values): conv1.W (C, nb_h + C, 3, 3), conv1.b, then for dilconv1..7
W (Cin, Cout, k, k) -- DilatedConv2DLayer stores (input, output) channels first -- and b.
"""
import os

import numpy as np
import torch

from .. import _kernels as K
from .._packing import load_npz_params
from .fcn8 import LayerHandle

DILATIONS = (1, 2, 4, 8, 16, 1)       # dilconv1..6, models/contextmod_dae.py:76-99
PAD = 32                              # PadLayer(width=32): sum of the dilated convs' valid shrink per side


class ContextModNet(object):
    split = False
    cm = 1
    fusable_update = False            # the loop runs its stand-alone softmax / update kernel on the logits
    mask_noise = 0.0
    takes_active = True               # logits(active=): images the early exit froze are skipped inside the kernels

    def __init__(self, n_classes, nb_h, params, device='cuda'):
        K.require_device()
        assert 1 <= n_classes <= 16 and 1 <= nb_h <= 16, 'context module: n_classes and the conditioning tensor have <= 16 channels'
        assert len(params) == 16, 'contextmod checkpoint: 8 layers x (W, b), got %d arrays' % len(params)
        self.n_classes, self.nb_h = n_classes, nb_h
        self.h_pad, self.y_cpad = nb_h, 16
        self.device = torch.device(device)
        p = [np.asarray(a, dtype=np.float32) for a in params]
        C_ = n_classes
        assert p[0].shape == (C_, nb_h + C_, 3, 3), ('conv1.W', p[0].shape)
        f = np.ascontiguousarray
        # conv1 = Conv2DLayer(flip_filters=False): W[f, c, r, s], the h channels first (models/model_helpers.py:91-93)
        self.w_h = f(p[0][:, :nb_h].transpose(1, 2, 3, 0))
        self.w_y = f(p[0][:, nb_h:].transpose(1, 2, 3, 0))
        self.b1 = f(p[1])
        self.zero_b = np.zeros((C_,), dtype=np.float32)
        self.dil = []
        for i in range(6):
            W, b = p[2 + 2 * i], p[3 + 2 * i]
            assert W.shape == (C_, C_, 3, 3), ('dilconv%d.W' % (i + 1), W.shape)
            self.dil.append((f(W.transpose(0, 2, 3, 1)), f(b)))        # (Cin, Cout, r, s) -> [Cin][r][s][Cout]
        assert p[14].shape == (C_, C_, 1, 1), ('dilconv7.W', p[14].shape)
        self.w7, self.b7 = f(p[14][:, :, 0, 0]), f(p[15])
        self._ws = {}

    def h_spatial(self, H, W):
        return H, W

    def alloc_h(self, B, H, W):
        return torch.zeros((B, self.nb_h, H, W), dtype=torch.float32, device=self.device)

    def pack_h(self, h, out=None):
        """h stays what the reference feeds: the NCHW float32 image."""
        assert h.dtype == torch.float32 and h.shape[1] == self.nb_h
        if out is None:
            return h.contiguous()
        out.copy_(h)
        return out

    def workspace(self, B, H, W):
        key = (B, H, W)
        ws = self._ws.get(key)
        if ws is None:
            C_, dev = self.n_classes, self.device
            c4 = (C_ + 3) & ~3          # the module's own tensors are channels-last, channels padded to whole 16-byte accesses
            n = B * c4 * (H + 2 * PAD) * (W + 2 * PAD)
            ws = {'padded': torch.zeros((B, H + 2 * PAD, W + 2 * PAD, c4), dtype=torch.float32, device=dev),   # border = PadLayer zeros
                  'ping': torch.empty((n,), dtype=torch.float32, device=dev),
                  'pong': torch.empty((n,), dtype=torch.float32, device=dev),
                  'hproj': torch.empty((B, H, W, c4), dtype=torch.float32, device=dev),
                  'logits': torch.empty((B, H, W, 16), dtype=torch.float32, device=dev)}
            self._ws[key] = ws
        return ws

    def executed_conv_flops(self, H, W, steady_state=True):
        """2*MAC per image of one application (the h half of conv1 only on the first iteration of a batch)."""
        C_ = self.n_classes
        fl = [2.0 * H * W * C_ * C_ * 9 + (0.0 if steady_state else 2.0 * H * W * self.nb_h * C_ * 9)]
        s = 2 * PAD
        for d in DILATIONS:
            s -= 2 * d
            fl.append(2.0 * (H + s) * (W + s) * C_ * C_ * 9)
        fl.append(2.0 * H * W * C_ * C_)
        return fl

    def logits(self, h, y_bf16=None, full_down=True, update=None, y_f32=None, noise=None, active=None):
        """h: NCHW fp32 (B, nb_h, H, W); y_f32: NCHW fp32 (B, C, H, W) -- the bf16 copy the tensor-core DAE reads is not
        used here.  `full_down`: h is new, recompute the hoisted W_h * h term.  Returns the fp32 NHWC16 logits (B, H, W, 16)."""
        assert update is None, 'context module: the softmax / update runs in the stand-alone kernel'
        assert y_f32 is not None and y_f32.dtype == torch.float32, 'context module reads the fp32 master y (y_f32=)'
        B, C_, H, W = y_f32.shape
        assert C_ == self.n_classes and tuple(h.shape) == (B, self.nb_h, H, W), (tuple(h.shape), tuple(y_f32.shape))
        ws = self.workspace(B, H, W)
        if full_down:
            K.ctx_conv(h, self.w_h, self.zero_b, 1, ws['hproj'], relu=False, origin=(-1, -1), check=True, out_nhwc=True)
        K.ctx_conv(y_f32, self.w_y, self.b1, 1, ws['padded'], relu=True, origin=(-1, -1), check=True, out_origin=(PAD, PAD),
                   size=(H, W), addend=ws['hproj'], active=active, out_nhwc=True)
        x, s = ws['padded'], 2 * PAD
        c4 = (C_ + 3) & ~3
        bufs = (ws['ping'], ws['pong'])
        for i, d in enumerate(DILATIONS):
            s -= 2 * d
            Wk, bk = self.dil[i]
            if i < 5:
                out = bufs[i & 1][:B * c4 * (H + s) * (W + s)].view(B, H + s, W + s, c4)
                K.ctx_conv(x, Wk, bk, d, out, relu=True, active=active, in_nhwc=True, out_nhwc=True)
                x = out
            else:
                assert s == 0
                K.ctx_conv(x, Wk, bk, d, ws['logits'], relu=True, tail=(self.w7, self.b7), active=active, in_nhwc=True)
        return ws['logits']


def buildDAE_contextmod(input_concat_h_vars, input_mask_var, n_classes, path_weights='/Tmp/romerosa/itinf/models/',
                        model_name='dae_model.npz', trainable=False, load_weights=False, out_nonlin=None,
                        concat_h=['input'], noise=0.1, params=None, nb_features_to_concat=3):
    """Same arguments as the reference builder (models/contextmod_dae.py:19-23); returns the handle of 'probs_dimshuffle'.
    The symbolic inputs are ignored; `noise` is GaussianNoiseLayer, the identity under deterministic=True
    (iterative_inference.py:189-190).  The conditioning tensor is the image: 3 channels in the reference
    (models/contextmod_dae.py:59), `nb_features_to_concat` for other inputs."""
    if not all(el in ['input'] for el in concat_h):
        raise AssertionError('context module does not reduce the image resolution: concat_h must be [\'input\']')   # :42
    if len(concat_h) != 1:
        raise NotImplementedError('B200 context module concatenates one conditioning tensor: concat_h=[\'input\']')
    if params is None:
        if not load_weights:
            raise ValueError('buildDAE_contextmod needs weights: pass params= or load_weights=True with path_weights')
        params = load_npz_params(os.path.join(path_weights, model_name))
    return LayerHandle(ContextModNet(n_classes, nb_features_to_concat, params), 'probs_dimshuffle', n_classes)

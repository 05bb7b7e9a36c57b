"""FCN8 segmentation net on the sm_100a kernels: drop-in for models/fcn8.py.

`buildFCN8` keeps the reference's signature (models/fcn8.py:16-22) and returns
one handle per name in `layer`; the handles stand in for the Lasagne layers the
reference returns (`fcn[0].output_shape[1]` is read at iterative_inference.py:152).
The forward pass is run by `FCN8Net.forward` -- 15 VGG/fc convolutions and three
1x1 score convolutions on the tcgen05 implicit-GEMM kernel, five max-pools, three
transposed convolutions and the softmax tail.
"""
import torch

from .. import _kernels as K
from .._packing import pack_conv, pack_deconv16, load_npz_params

VGG_STAGES = [  # models/fcn8.py:33-72
    [('conv1_1', 64), ('conv1_2', 64)],
    [('conv2_1', 128), ('conv2_2', 128)],
    [('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256)],
    [('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512)],
    [('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512)],
]
PARAM_ORDER = [n for st in VGG_STAGES for n, _ in st] + [
    'fc6', 'fc7', 'score_fr', 'score2', 'score_pool4', 'score4', 'score_pool3', 'upsample']
_POOL_CHANNELS = {'pool1': 64, 'pool2': 128, 'pool3': 256, 'pool4': 512, 'pool5': 512}


class LayerHandle(object):
    """Stands in for a Lasagne layer: carries the name, the symbolic output shape
    (None for batch / spatial, like InputLayer((None, C, None, None))) and the net."""

    def __init__(self, net, name, channels):
        self.net = net
        self.name = name
        self.output_shape = (None, channels, None, None)

    def __repr__(self):
        return 'LayerHandle(%s, %r)' % (self.name, self.output_shape)


class FCN8Net(object):
    def __init__(self, nb_in_channels, n_classes, params, temperature=1.0, device='cuda', precision='bf16', concat=None):
        """precision: 'bf16', or 'fp32x3' / 'mixed' (see DAENet; FCN8 has no expanding path of its own and its output
        y0 enters y directly, so both run every layer fp32-accurate).
        `concat` = (layer, nb_h), layer in 'input', 'pool1'..'pool4': the FCN8-shaped DAE (models/fcn8_dae.py:46-115) --
        nb_h conditioning channels are concatenated in front of that layer's output, so the next conv has nb_h more input
        channels.  Its h half is iteration-invariant: it is packed as a separate bias-free conv whose fp32 result
        (`forward(h=...)`, once per batch) the conv on the other half adds in its epilogue."""
        K.require_device()
        assert precision in ('bf16', 'fp32x3', 'mixed'), precision
        self.precision = precision
        self.split = sp = precision != 'bf16'
        self.cm = 2 if sp else 1
        assert n_classes <= 16
        assert len(params) == 2 * len(PARAM_ORDER), 'expected %d arrays, got %d' % (2 * len(PARAM_ORDER), len(params))
        self.nb_in_channels = nb_in_channels
        self.n_classes = n_classes
        self.device = torch.device(device)
        P = {n: (params[2 * i], params[2 * i + 1]) for i, n in enumerate(PARAM_ORDER)}
        self.w = {}
        cin = nb_in_channels
        self.concat, self.concat_conv, self.hproj = concat, None, None
        if concat is not None:
            assert concat[0] in ('input', 'pool1', 'pool2', 'pool3', 'pool4'), concat
        pending = concat is not None and concat[0] == 'input'
        for si, stage in enumerate(VGG_STAGES):
            for name, cout in stage:
                Wn, bn = P[name]
                if pending:        # W[:, :nb_h] acts on h (ConcatLayer((h, layer)), models/model_helpers.py:91-93)
                    nb_h = concat[1]
                    assert Wn.shape[1] == nb_h + cin, (name, tuple(Wn.shape), nb_h, cin)
                    self.concat_conv = name
                    self.h_pad = K.pad_channels(nb_h)        # (the fp32-out hoisted conv runs on the 64-channel-block kernels)
                    self.w['hproj'] = pack_conv(Wn[:, :nb_h], torch.zeros_like(torch.as_tensor(bn)), [(nb_h, self.h_pad)], cout,
                                                self.device, split=sp)
                    Wn = Wn[:, nb_h:]
                    pending = False
                cpad = K.pad_channels(cin, narrow=(name == 'conv1_1'))
                self.w[name] = pack_conv(Wn, bn, [(cin, cpad)], cout, self.device, split=sp)
                cin = cout
            if concat is not None and concat[0] == 'pool%d' % (si + 1):
                pending = True
        self.w['fc6'] = pack_conv(*P['fc6'], [(512, 512)], 4096, self.device, split=sp)
        self.w['fc7'] = pack_conv(*P['fc7'], [(4096, 4096)], 4096, self.device, split=sp)
        self.w['score_fr'] = pack_conv(*P['score_fr'], [(4096, 4096)], 16, self.device, split=sp)
        self.w['score_pool4'] = pack_conv(*P['score_pool4'], [(512, 512)], 16, self.device, split=sp)
        self.w['score_pool3'] = pack_conv(*P['score_pool3'], [(256, 256)], 16, self.device, split=sp)
        self.w['score2'] = pack_deconv16(*P['score2'], self.device)
        self.w['score4'] = pack_deconv16(*P['score4'], self.device)
        # temperature divides upsample.W and .b (models/fcn8.py:193-198)
        self.w['upsample'] = pack_deconv16(*P['upsample'], self.device, scale=1.0 / float(temperature))

    def forward(self, X, want=('pool4', 'probs_dimshuffle'), y_bf16_cpad=None, x_packed=None, h=None):
        """X: NCHW fp32 CUDA (B, nb_in_channels, H, W).  Returns a dict with the
        requested names: 'poolK' -> NHWC bf16, 'probs_dimshuffle' -> NCHW fp32, 'logits' -> the fp32 NHWC16 rows in front
        of the softmax, plus 'y_bf16' (NHWC bf16, `y_bf16_cpad` channels) when requested.
        `x_packed`: the input already as NHWC bf16 (pairs), (B, H, W, cm*16), instead of X.
        `h` (nets built with concat=): the packed conditioning tensor (NHWC bf16 (pairs), h_pad channels); given once per
        batch, it refreshes the hoisted fp32 term the concat conv adds, which is kept for the following calls."""
        sp, cm = self.split, self.cm
        out = {}
        if x_packed is not None:
            B, H, W, _ = x_packed.shape
            assert x_packed.shape[3] == cm * K.pad_channels(self.nb_in_channels, narrow=True)
            x = x_packed
        else:
            B, Cin, H, W = X.shape
            assert Cin == self.nb_in_channels
            if 'input' in want:          # net['input'] (models/fcn8.py:30): the image itself, the conditioning of concat_h=['input']
                out['input'] = X
            x = K.pack_nchw(X.contiguous(), K.pad_channels(Cin, narrow=True), split=sp)
        dev = x.device
        if self.concat is not None:
            if h is not None:
                pad = 100 if self.concat_conv == 'conv1_1' else 1
                self.hproj = K.conv2d(h, *self.w['hproj'], 3, 3, pad, relu=False, out_f32=True, split=sp)
            assert self.hproj is not None, 'FCN8 with a concatenated input: pass h= on the first call of a batch'
        for si, stage in enumerate(VGG_STAGES):
            for ci, (name, cout) in enumerate(stage):
                Wk, bk = self.w[name]
                pad = 100 if name == 'conv1_1' else 1
                kw = dict(addend=self.hproj) if name == self.concat_conv else {}
                if ci == len(stage) - 1:    # last conv of the stage: max-pool fused in the epilogue
                    oh, ow = K.conv_out_size(x.shape[1], x.shape[2], 3, 3, pad)
                    pooled = torch.empty((B, oh // 2, ow // 2, cm * cout), dtype=torch.bfloat16, device=dev)
                    x = K.conv2d(x, Wk, bk, 3, 3, pad, relu=True, pooled=pooled, split=sp)
                else:
                    x = K.conv2d(x, Wk, bk, 3, 3, pad, relu=True, split=sp, **kw)
            out['pool%d' % (si + 1)] = x
        x = K.conv2d(x, *self.w['fc6'], 7, 7, 0, relu=True, split=sp)
        x = K.conv2d(x, *self.w['fc7'], 1, 1, 0, relu=True, split=sp)
        score_fr = K.conv2d(x, *self.w['score_fr'], 1, 1, 0, relu=True, out_f32=True, split=sp)
        # score2 + centre-cropped score_pool4 (models/fcn8.py:90-97)
        sp4 = K.conv2d(out['pool4'], *self.w['score_pool4'], 1, 1, 0, relu=True, out_f32=True, split=sp)
        fused = self._deconv_sum(score_fr, 'score2', 4, 2, sp4)
        sp3 = K.conv2d(out['pool3'], *self.w['score_pool3'], 1, 1, 0, relu=True, out_f32=True, split=sp)
        final = self._deconv_sum(fused, 'score4', 4, 2, sp3)
        # upsample + centre crop to the input size (models/fcn8.py:109-118)
        fH, fW = (final.shape[1] - 1) * 8 + 16, (final.shape[2] - 1) * 8 + 16
        assert fH >= H and fW >= W
        logits = K.deconv16(final, *self.w['upsample'], 16, 8, window=((fH - H) // 2, (fW - W) // 2, H, W))
        if 'logits' in want:
            out['logits'] = logits
            if 'probs_dimshuffle' not in want:
                return {k: v for k, v in out.items() if k in want}
        probs = torch.empty((B, self.n_classes, H, W), dtype=torch.float32, device=dev)
        y_bf16 = None
        if y_bf16_cpad:
            y_bf16 = torch.empty((B, H, W, cm * y_bf16_cpad), dtype=torch.bfloat16, device=dev)
        K.softmax_nchw(logits, self.n_classes, probs, y_bf16, split=sp)
        out['probs_dimshuffle'] = probs
        out['y_bf16'] = y_bf16
        return {k: v for k, v in out.items() if k in want or k == 'y_bf16'}

    def _deconv_sum(self, x, name, k, stride, other):
        """Deconv then ElemwiseSumLayer(cropping='center') with `other`: both are
        centre-cropped to the per-axis minimum size."""
        fH, fW = (x.shape[1] - 1) * stride + k, (x.shape[2] - 1) * stride + k
        mh, mw = min(fH, other.shape[1]), min(fW, other.shape[2])
        return K.deconv16(x, *self.w[name], k, stride, window=((fH - mh) // 2, (fW - mw) // 2, mh, mw),
                          addend=other, addend_off=((other.shape[1] - mh) // 2, (other.shape[2] - mw) // 2))


def buildFCN8(nb_in_channels, input_var=None,
              path_weights='/Tmp/romerosa/itinf/models/camvid/new_fcn8_model_best.npz',
              n_classes=21, load_weights=True, void_labels=[], trainable=False,
              layer=['probs_dimshuffle'], pascal=False, temperature=1.0, dropout=0.5,
              params=None, precision='bf16'):
    """Same arguments as the reference builder (models/fcn8.py:16-22).  `input_var` is a
    Theano symbol there and is ignored here; `params` (a 42-array list in checkpoint
    order) may be passed instead of `path_weights`.  Inference only: `trainable` and
    `dropout` have no effect on the deterministic forward pass, `pascal` (.mat weights)
    is not supported."""
    if pascal:
        raise NotImplementedError('pascal .mat weights are outside the iterative-inference path')
    if params is None:
        if not load_weights:
            raise ValueError('buildFCN8 needs weights: pass params= or load_weights=True with path_weights')
        params = load_npz_params(path_weights)
    # NB the reference applies `temperature` only when load_weights is set (models/fcn8.py:194)
    net = FCN8Net(nb_in_channels, n_classes, params, temperature=temperature if load_weights else 1.0,
                  precision=precision)
    handles = []
    for el in layer:
        if el in _POOL_CHANNELS:
            handles.append(LayerHandle(net, el, _POOL_CHANNELS[el]))
        elif el == 'input':
            handles.append(LayerHandle(net, el, nb_in_channels))
        elif el == 'probs_dimshuffle':
            handles.append(LayerHandle(net, el, n_classes))
        else:
            raise ValueError('layer %r is not exposed by the B200 FCN8 (input, pool1..pool5, probs_dimshuffle)' % el)
    return handles

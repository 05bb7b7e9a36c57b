"""FC-DenseNet103 segmentation net on the sm_100a kernels: drop-in for models/FCDenseNet.py.

`build_fcdensenet` keeps the reference's signature (models/FCDenseNet.py:196-219) and returns
`hidden_outputs + [output]` as handles (`fcn[0].output_shape[1]` is read at
iterative_inference.py:152).  The network is models/FCDenseNet.py:61-146 with the layer helpers of the
un-vendored `FC_DenseNet.layers` (SimJeg/FC-DenseNet; BN_ReLU_Conv, TransitionDown, TransitionUp,
SoftmaxLayer -- semantics restated in oracle/densenet.py):

  * the Tiramisu stack of a dense block is ONE fp32 NHWC tensor allocated with its final channel
    count; a layer's 16 new feature maps are written by the conv epilogue as a channel slice behind
    the existing ones (`out_slice`), so no ConcatLayer ever copies the stack;
  * BN_ReLU_Conv = iiseg_bn_relu_pack (BatchNorm with BATCH statistics, iterative_inference.py:187
    batch_norm_use_averages=False, + rectify + bf16 pack) -> tcgen05 implicit-GEMM conv (Cout = 16,
    fp32 output slice) -> iiseg_channel_stats of the 16 new maps (statistics of a feature map do not
    depend on the consuming layer, so each map is reduced exactly once);
  * TransitionDown = BN/ReLU pack -> 1x1 conv (fp32 out) -> iiseg_maxpool2_f32 into the next stack;
  * TransitionUp = Deconv2DLayer(3, stride 2, crop 'valid') as its four output-phase convolutions
    (2x2, 2x1, 1x2 and 1x1 taps of the flipped kernel) on the tensor cores + iiseg_deconv_interleave
    with the centre crop, deconv maps first, then the skip stack (ConcatLayer([l, skip]));
  * SoftmaxLayer = 1x1 conv to fp32 logits -> iiseg_softmax_nchw.

bf16 operands / fp32 accumulation and fp32 stacks (precision 'bf16'); the fp32x3 variant is not built
for this net.  Results depend on the batch composition (batch statistics): shard whole batches.
"""
import torch

from .. import _kernels as K
from .._packing import pack_conv, load_npz_params, _as_f32
from .fcn8 import LayerHandle

N_LAYERS_103 = [4, 5, 7, 10, 12, 15, 12, 10, 7, 5, 4]      # models/FCDenseNet.py:205
BN_EPS = 1e-4                                             # lasagne BatchNormLayer default


def _r64(c):
    return (c + 63) // 64 * 64


class _Stack(object):
    """fp32 NHWC feature stack + the batch statistics of the channels produced so far."""

    def __init__(self, B, H, W, C, dev):
        self.t = torch.empty((B, H, W, C), dtype=torch.float32, device=dev)
        self.mean = torch.zeros((C,), dtype=torch.float32, device=dev)
        self.inv_std = torch.ones((C,), dtype=torch.float32, device=dev)
        self.n = 0      # channels filled


class DenseNetNet(object):
    def __init__(self, nb_in_channels, n_classes, params, n_first=48, n_pool=5, growth=16,
                 n_layers=N_LAYERS_103, device='cuda', precision='bf16'):
        """precision: 'bf16' (bf16 conv operands, fp32 stacks / statistics / accumulation) or 'fp32x3' / 'mixed' (the net
        has no expanding path of the DAE kind, so both mean: every conv fp32-accurate -- BN + rectify packs the (hi, lo)
        bf16 pair of its fp32 result and each conv accumulates hi*hi + lo*hi + hi*lo; see DAENet)."""
        K.require_device()
        assert precision in ('bf16', 'fp32x3', 'mixed'), precision
        self.precision = precision
        assert n_classes <= 16 and nb_in_channels <= 16 and growth == 16
        self.nb_in, self.n_classes = nb_in_channels, n_classes
        self.n_first, self.n_pool, self.growth, self.n_layers = n_first, n_pool, growth, list(n_layers)
        self.device = dev = torch.device(device)
        self.split = sp = precision != 'bf16'     # interface shared with FCN8Net
        self.cm = 2 if sp else 1
        it = iter(params)

        def vec(a):
            return _as_f32(a, dev).contiguous()

        def bnconv(cin, cout_pad, k):
            beta, gamma, _mean, _inv_std = next(it), next(it), next(it), next(it)   # stored averages are never read (:187)
            W, b = next(it), next(it)
            assert tuple(W.shape[1:]) == (cin, k, k), (tuple(W.shape), cin, k)
            Wk, bk = pack_conv(W, b, [(cin, _r64(cin))], cout_pad, dev, split=sp)
            return {'gamma': vec(gamma), 'beta': vec(beta), 'W': Wk, 'b': bk, 'cin': cin}

        W, b = next(it), next(it)
        self.first = pack_conv(W, b, [(nb_in_channels, 64 if sp else 16)], _r64(n_first), dev, split=sp)
        n = n_first
        self.down, self.td, self.skip_ch = [], [], []
        for i in range(n_pool):
            blk = []
            for j in range(n_layers[i]):
                blk.append(bnconv(n, 16, 3))
                n += growth
            self.down.append(blk)
            self.skip_ch.append(n)
            self.td.append(bnconv(n, _r64(n), 1))
        self.bottleneck = []
        self.bott_in = n
        for j in range(n_layers[n_pool]):
            self.bottleneck.append(bnconv(n, 16, 3))
            n += growth
        self.tu, self.up = [], []
        up_ch = growth * n_layers[n_pool]
        for i in range(n_pool):
            keep = growth * n_layers[n_pool + i]
            W, b = next(it), next(it)
            assert tuple(W.shape) == (up_ch, keep, 3, 3), (tuple(W.shape), up_ch, keep)
            self.tu.append(self._pack_deconv(W, b, up_ch, keep))
            n = keep + self.skip_ch[n_pool - 1 - i]
            blk = []
            for j in range(n_layers[n_pool + i + 1]):
                blk.append(bnconv(n, 16, 3))
                n += growth
            self.up.append(blk)
            up_ch = growth * n_layers[n_pool + i + 1]
        W, b = next(it), next(it)
        assert tuple(W.shape) == (n_classes, n, 1, 1)
        self.final = pack_conv(W, b, [(n, _r64(n))], 16, dev, split=sp)
        self.final_in = n
        assert next(it, None) is None, 'unused parameters'
        self._ws = {}
        self._graphs = {}

    def _pack_deconv(self, W, b, cin, keep):
        """Deconv2DLayer W (in, out, 3, 3), flip_filters=False = conv_transpose2d with the flipped kernel Wf.
        Output phase (py, px): out[2i+py, 2j+px] = sum over taps a = oy - 2*iy in {py, py+2} (< 3), i.e. a
        stride-1 cross-correlation with R = 2 - py taps: pad 1, tap r reads x[i + r - 1] -> a = (2, 0)[r] for
        py = 0 and a = 1 for py = 1 (R = 1; the launch shifts its output window by one instead of padding 0)."""
        Wf = _as_f32(W, self.device).flip(2, 3)
        phases = []
        for py in range(2):
            row = []
            for px in range(2):
                ah = [2, 0] if py == 0 else [1]
                aw = [2, 0] if px == 0 else [1]
                Wc = Wf[:, :, ah][:, :, :, aw].permute(1, 0, 2, 3).contiguous()      # (out, in, R, S)
                row.append(pack_conv(Wc, b, [(cin, _r64(cin))], _r64(keep), self.device, split=getattr(self, 'split', False)) + (len(ah), len(aw)))
            phases.append(row)
        return {'phases': phases, 'cin': cin, 'keep': keep}

    # -- building blocks -----------------------------------------------------
    def _scratch(self, B, H, W, C):
        need = K._lib.load().iiseg_channel_stats_chunks(B, H, W) * C * 2
        s = self._ws.get('scratch')
        if s is None or s.numel() < need:
            s = self._ws['scratch'] = torch.empty((need,), dtype=torch.float64, device=self.device)
        return s

    def _packbuf(self, B, H, W, C):
        key = ('pack', B, H, W, C)
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = torch.empty((B, H, W, self.cm * C), dtype=torch.bfloat16, device=self.device)
        return t

    def _stats(self, st, c0, C):
        B, H, W, _ = st.t.shape
        K.channel_stats(st.t, c0, C, st.mean, st.inv_std, self._scratch(B, H, W, C), eps=BN_EPS)

    def _dense_layer(self, st, lay):
        """BN_ReLU_Conv(stack, 16) appended to the stack (models/FCDenseNet.py:84-89)."""
        B, H, W, _ = st.t.shape
        C = lay['cin']
        assert st.n == C
        xb = K.bn_relu_pack(st.t, C, self._packbuf(B, H, W, _r64(C)), stats=(st.mean, st.inv_std), gamma=lay['gamma'],
                            beta=lay['beta'], relu=True, split=self.split)
        K.conv2d(xb, lay['W'], lay['b'], 3, 3, 1, relu=False, out_f32=True, out_slice=(st.t, C), split=self.split)
        self._stats(st, C, 16)
        st.n = C + 16

    def forward(self, X, want=('pool4', 'probs_dimshuffle'), y_bf16_cpad=None, use_graph=True):
        """The ~700 launches of the forward pass are captured once per (input shape, `want`) as a CUDA graph and
        replayed (eagerly the pass is bound by the host's launch rate: 18-30 ms of host time for 18 ms of kernels);
        the results are copied out of the graph's static buffers, so every call returns fresh tensors."""
        if not use_graph:
            return self._forward_eager(X, want, y_bf16_cpad)
        key = (tuple(X.shape), tuple(want), y_bf16_cpad)
        ent = self._graphs.get(key)
        if ent is None:
            xs = X.contiguous().clone()
            self._forward_eager(xs, want, y_bf16_cpad)          # warm-up: workspaces, kernel attributes
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = self._forward_eager(xs, want, y_bf16_cpad)
            ent = self._graphs[key] = (g, xs, outs)
        g, xs, outs = ent
        xs.copy_(X)
        g.replay()
        return {k: (v.clone() if v is not None else None) for k, v in outs.items()}

    def _forward_eager(self, X, want=('pool4', 'probs_dimshuffle'), y_bf16_cpad=None):
        """X: NCHW fp32 CUDA.  Returns {'poolK': fp32 NHWC stack after the K-th TransitionDown, 'poolK_bf16': its
        64-padded bf16 copy (the DAE's conditioning input), 'probs_dimshuffle': NCHW fp32, 'y_bf16': optional
        NHWC bf16 copy of the probabilities}."""
        B, Cin, H, W = X.shape
        assert Cin == self.nb_in
        dev, g = X.device, self.growth
        out = {}
        sp = self.split
        x16 = K.pack_nchw(X.contiguous(), 64 if sp else 16, split=sp)
        st = _Stack(B, H, W, self.skip_ch[0], dev)
        first = K.conv2d(x16, self.first[0], self.first[1], 3, 3, 1, relu=False, out_f32=True, split=sp)     # 64-padded fp32
        st.t[..., :self.n_first].copy_(first[..., :self.n_first])
        st.n = self.n_first
        self._stats(st, 0, self.n_first)
        skips = []
        h, w = H, W
        for i in range(self.n_pool):
            for lay in self.down[i]:
                self._dense_layer(st, lay)
            skips.append(st)
            # TransitionDown: BN_ReLU_Conv(stack, n, 1x1) -> maxpool 2
            td, C = self.td[i], st.n
            xb = K.bn_relu_pack(st.t, C, self._packbuf(B, h, w, _r64(C)), stats=(st.mean, st.inv_std), gamma=td['gamma'],
                                beta=td['beta'], relu=True, split=sp)
            y = K.conv2d(xb, td['W'], td['b'], 1, 1, 0, relu=False, out_f32=True, split=sp)
            h, w = h // 2, w // 2
            n_next = self.skip_ch[i + 1] if i + 1 < self.n_pool else self.bott_in + g * self.n_layers[self.n_pool]
            nxt = _Stack(B, h, w, n_next, dev)
            K.maxpool2_f32(y, C, nxt.t)
            nxt.n = C
            self._stats(nxt, 0, C)
            st = nxt
            name = 'pool%d' % (i + 1)
            if name in want:      # the stack right after the transition (models/FCDenseNet.py:96-97)
                out[name] = st.t[..., :C].contiguous()                                   # fp32 NHWC [B,h,w,C]
                out[name + '_bf16'] = K.bn_relu_pack(st.t, C, torch.empty((B, h, w, self.cm * _r64(C)), dtype=torch.bfloat16, device=dev),
                                                     relu=False, split=sp)                # what DAENet.logits reads
        for lay in self.bottleneck:
            self._dense_layer(st, lay)
        up_c0, up_ch = self.bott_in, g * self.n_layers[self.n_pool]       # block_to_upsample = the block's new maps
        for i in range(self.n_pool):
            tu, skip = self.tu[i], skips[self.n_pool - 1 - i]
            keep = tu['keep']
            # Deconv2DLayer(3, stride 2) of the (un-normalised) block maps, as four phase convolutions
            xb = K.bn_relu_pack(st.t, up_ch, self._packbuf(B, h, w, _r64(up_ch)), c0=up_c0, relu=False, split=sp)
            phases = []
            for py in range(2):
                row = []
                for px in range(2):
                    Wk, bk, R, S = tu['phases'][py][px]
                    row.append(K.conv2d(xb, Wk, bk, R, S, 1, relu=False, out_f32=True, split=sp,
                                        window=(1 if R == 1 else 0, 1 if S == 1 else 0, h + 1, w + 1)))
                phases.append(row)
            sh, sw = skip.t.shape[1], skip.t.shape[2]
            dh, dw = 2 * h + 1, 2 * w + 1
            mh, mw = min(dh, sh), min(dw, sw)              # ConcatLayer(cropping=center): per-axis minimum
            assert (mh, mw) == (sh, sw), 'skip maps are never larger than the deconv output here'
            n_blk = self.n_layers[self.n_pool + i + 1]
            nxt = _Stack(B, mh, mw, keep + skip.n + g * n_blk, dev)
            K.deconv_interleave(phases, keep, ((dh - mh) // 2, (dw - mw) // 2), nxt.t)
            self._stats(nxt, 0, keep)
            nxt.t[..., keep:keep + skip.n].copy_(skip.t[..., :skip.n])           # the skip stack (data movement only)
            nxt.mean[keep:keep + skip.n].copy_(skip.mean[:skip.n])
            nxt.inv_std[keep:keep + skip.n].copy_(skip.inv_std[:skip.n])
            nxt.n = keep + skip.n
            st, h, w = nxt, mh, mw
            up_c0, up_ch = st.n, g * n_blk
            for lay in self.up[i]:
                self._dense_layer(st, lay)
        assert st.n == self.final_in and (h, w) == (H, W)
        xb = K.bn_relu_pack(st.t, st.n, self._packbuf(B, h, w, _r64(st.n)), relu=False, split=sp)          # SoftmaxLayer: no BN
        logits = K.conv2d(xb, self.final[0], self.final[1], 1, 1, 0, relu=False, out_f32=True, split=sp)
        probs = torch.empty((B, self.n_classes, H, W), dtype=torch.float32, device=dev)
        y_bf16 = None
        if y_bf16_cpad:
            y_bf16 = torch.empty((B, H, W, self.cm * y_bf16_cpad), dtype=torch.bfloat16, device=dev)
        K.softmax_nchw(logits, self.n_classes, probs, y_bf16, split=sp)
        out['probs_dimshuffle'] = probs
        out['y_bf16'] = y_bf16
        return out


_POOL_CHANNELS_103 = {'pool1': 112, 'pool2': 192, 'pool3': 304, 'pool4': 464, 'pool5': 656}


def build_fcdensenet(input_var, layer, nb_in_channels=3, n_classes=11, output_d='4d', from_gt=False,
                     weight_path='/data/lisatmp4/romerosa/itinf/models/camvid/DenseNet103/weights/FC-DenseNet103_weights.npz',
                     params=None, precision='bf16'):
    """Same arguments as the reference builder (models/FCDenseNet.py:196-198); `input_var` (a Theano
    symbol there) is ignored, `params` (the positional checkpoint arrays) may replace `weight_path`.
    Returns hidden_outputs (one handle per 'poolK' in `layer`) + [output] unless from_gt."""
    if output_d != '4d':
        raise NotImplementedError("output_d='2d' is not used on the iterative-inference path")
    if params is None:
        params = load_npz_params(weight_path)
    net = DenseNetNet(nb_in_channels, n_classes, params, precision=precision)
    handles = []
    for el in layer:
        if el not in _POOL_CHANNELS_103:
            raise ValueError('layer %r is not exposed by the B200 FC-DenseNet103 (pool1..pool5)' % el)
        handles.append(LayerHandle(net, el, _POOL_CHANNELS_103[el]))
    if not from_gt:
        handles.append(LayerHandle(net, 'probs_dimshuffle', n_classes))
    return handles

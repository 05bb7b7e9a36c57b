"""DAE_h (conditional denoising auto-encoder) on the sm_100a kernels: drop-in for
models/DAE_h.py + models/fcn_down.py + models/fcn_up.py + layers/mylayers.py.

One application DAE(y, h), benchmark configuration (kind='standard',
unpool_type='trackind', conv_before_pool=1, skip=True, bn=0, dropout=0):

  down, p = 1..P : conv3x3+ReLU (tcgen05 implicit GEMM; pad=`padding` on level 1,
                   the (h, pool_n) concat of level n_pool+1 read by the loader from
                   two tensor maps) with the 2x2 max-pool + tie-inclusive mask fused
                   in the epilogue (the pre-pool map never reaches HBM)
  up,   p = P..1 : mask unpool (DePool2D) -> conv3x3 linear with the skip-sum of
                   pool_{p-1} fused in the epilogue; level 1 computes only the
                   centre-crop window and emits fp32 logits
  tail           : channel softmax (+ the y update) in one streaming kernel
"""
import os

import torch

from .. import _kernels as K
from .._packing import pack_conv, pack_npack16, load_npz_params
from .fcn8 import LayerHandle


def _levels(concat_h, additional_pool):
    last = concat_h[-1]
    n_pool = int(last[-1]) if 'pool' in last else 0   # models/DAE_h.py:36-39
    return n_pool, n_pool + additional_pool


class DAENet(object):
    def __init__(self, n_classes, nb_features_to_concat, padding, params, concat_h=('pool4',),
                 n_filters=64, additional_pool=2, device='cuda', precision='bf16', unpool_type='trackind', bn=False,
                 mask_noise=0.0, skip=True, conv_before_pool=1):
        """precision: 'bf16' (bf16 operands, fp32 accumulation: the throughput variant), 'fp32x3'
        (every activation and weight is a (hi, lo) bf16 pair and each conv accumulates
        hi*hi + lo*hi + hi*lo on the same tensor-core loop: fp32-accurate, ~3x the MMA work) or
        'mixed': fp32x3 on the contracting path -- whose conv outputs decide the pool tie masks, the one
        discontinuous function on the path -- and bf16 on the expanding path, which only carries smooth
        errors (oracle/precision_mix.py: same parity numbers as 'fp32x3' on every iteration)."""
        K.require_device()
        assert precision in ('bf16', 'fp32x3', 'mixed'), precision
        # 'inverse' (lasagne InverseLayer of the pool, models/fcn_up.py:76-79) IS DePool2D's repeat * tie-mask under Theano's
        # CPU MaxPoolGrad (the gradient of a max-pool w.r.t. its input goes to every tied maximum); 'standard'
        # (models/fcn_up.py:37-63) replaces unpool + conv by Deconv2DLayer(4, stride=2), run as four 2x2 phase convs
        assert unpool_type in ('trackind', 'inverse', 'standard'), unpool_type
        self.unpool_type = unpool_type
        # mask_noise > 0 (`buildDAE(..., noise > 0)`; `stochastic_masks=False` turns it off): the reference's DePool2D builds its tie masks with
        # lasagne.layers.get_output(...) WITHOUT deterministic=True (layers/mylayers.py:91-93), so when the DAE was built with
        # noise > 0 the masks come from SEPARATE passes of the contracting path on y + N(0, noise^2), even at inference: one
        # per DePool2D, each with its own draw (see `logits`).  In this mode every application runs the value pass on y plus,
        # for level p, a mask pass over levels 1..p.
        self.mask_noise = float(mask_noise) if unpool_type == 'trackind' else 0.0          # InverseLayer / Deconv2DLayer take the deterministic expressions
        # skip=False (models/fcn_up.py:103-113): no ElemwiseSumLayer; the up-conv is only centre-cropped to the size of
        # pool_{p-1} (CroppingLayer with merge_function = lambda input, deconv: deconv) -- the same windows, no addend
        self.skip = bool(skip)
        self.precision = precision
        self.split = precision != 'bf16'          # contracting path (and the h / y input format)
        self.split_up = precision == 'fp32x3'     # expanding path
        # IISEG_FUSE_DEPOOL=1: last DePool2D expanded inside the loader of up_conv1 (bf16 variant; bit-identical results).
        # Off by default: measured 0.179 ms against 0.050 (unpool) + 0.112 (conv) -- see DESIGN.md 3.7
        self.fuse_depool = os.environ.get('IISEG_FUSE_DEPOOL', '0') == '1'
        # IISEG_DEPOOL_EPILOGUE=1: DePool2D of level p-1 written by the epilogue of up_conv_p (bf16 variant, bit-identical).
        # Off by default: measured 163 vs 165 images/s -- the 4x larger epilogue stores cost what the unpool launch did
        self.fuse_depool_out = os.environ.get('IISEG_DEPOOL_EPILOGUE', '0') == '1'
        self.cm = 2 if self.split else 1          # bf16 channels per logical channel in activation tensors (contracting path)
        self.cm_up = 2 if self.split_up else 1
        assert n_classes <= 16
        self.n_classes = n_classes
        self.nb_h = nb_features_to_concat
        self.h_pad = K.pad_channels(nb_features_to_concat)
        # concat_h = ['poolN']: h joins the DAE's own pool_N (n_pool = N); concat_h = ['input']: h (the image, 'input' layer of
        # the segmentation net) joins y at the DAE's input (n_pool = 0, models/model_helpers.py:86-96) and -- concatenation
        # only at the input -- the first conv is not padded by `padding` but 'same' (models/fcn_down.py:90-95)
        self.n_pool, self.total = _levels(concat_h, additional_pool)
        assert self.total >= 1, 'It seems your DAE will have no conv/pooling layers!'      # models/fcn_down.py:73-74
        if self.n_pool == 0:
            padding = 0
        self.padding = padding
        self.device = torch.device(device)
        self.y_cpad = K.pad_channels(n_classes, narrow=True)    # channels of the bf16 copy of y (16: 32-byte K blocks)
        # bn=1 (models/fcn_down.py:113-115, models/fcn_up.py:91-93): a BatchNormLayer behind every conv.  At inference the graph
        # is built with deterministic=True (iterative_inference.py:189-190), so the layer is the affine map
        # (x - mean) * (gamma * inv_std) + beta on its stored averages (checkpoint order beta, gamma, mean, inv_std):
        #   contracting path: conv -> rectify -> BN -> pool  = a per-channel scale / shift in the conv epilogue, after the
        #                     rectifier and before the pool + tie mask (iiseg_conv_desc.post_scale);
        #   expanding path  : conv (linear) -> BN -> skip-sum = folded into the conv's weights and bias when packing.
        # conv_before_pool = k > 1 (models/fcn_down.py:83-115): k rectified 3x3 convs per level, 'same' padding after the first;
        # the pool (and its tie mask) follows the last one.  The extra convs run on full maps every iteration (no y-dependent
        # windows: the region that depends on y then grows faster on the way down than the crop cone does on the way up).
        self.cbp = int(conv_before_pool)
        assert self.cbp >= 1 and not (self.cbp > 1 and bn), 'conv_before_pool > 1 is built for bn=0'
        self.down_extra = [[] for _ in range(self.total)]
        if self.cbp > 1:
            k = self.cbp
            assert len(params) == 2 * (k + 1) * self.total, 'conv_before_pool=%d: expected %d arrays, got %d' % (k, 2 * (k + 1) * self.total, len(params))
            first, extra = [], []
            for lvl in range(self.total):
                grp = params[2 * k * lvl:2 * k * (lvl + 1)]
                first += grp[:2]
                extra.append(grp[2:])
            params = first + list(params[2 * k * self.total:])
        self.bn = bool(bn)
        self.post = [None] * self.total
        self.bn_gb = [None] * self.total
        # bn=1 with DePool2D: the reference builds the tie masks in a sub-graph of its own, lasagne.layers.get_output(...)
        # WITHOUT deterministic=True (layers/mylayers.py:91-93), where every BatchNormLayer normalises with the statistics of
        # the CURRENT BATCH instead of its stored averages (lasagne: batch_norm_use_averages defaults to `deterministic`).
        # Observed by executing the reference (tests/golden/ref_bn.npz).  Every application therefore runs the contracting
        # path twice: the mask pass (conv -> fp32 map -> iiseg_channel_stats -> conv again with the batch-statistics affine,
        # pool and tie mask in its epilogue) and the value pass on the stored averages.  'inverse' (InverseLayer) and
        # 'standard' take the deterministic expressions and need no second pass.
        self.bn_batch_masks = self.bn and unpool_type == 'trackind'
        if self.bn:
            from .._packing import _as_f32
            k_up = 2 if unpool_type == 'standard' else 6
            assert len(params) == (6 + k_up) * self.total, 'bn=1: expected %d arrays, got %d' % ((6 + k_up) * self.total, len(params))
            flat = []
            for i in range(self.total):
                W, b, beta, gamma, mean, inv_std = [_as_f32(a, self.device) for a in params[6 * i:6 * i + 6]]
                s_ = gamma * inv_std
                # (scale, shift, mean): the epilogue evaluates ((x - mean) * (gamma * inv_std)) + beta with lasagne's roundings
                self.post[i] = (s_.contiguous(), beta.contiguous(), mean.contiguous())
                self.bn_gb[i] = (beta.contiguous(), gamma.contiguous())
                flat += [W, b]
            for i in range(self.total):
                grp = [_as_f32(a, self.device) for a in params[6 * self.total + k_up * i:6 * self.total + k_up * (i + 1)]]
                W, b = grp[0], grp[1]
                if k_up == 6:
                    beta, gamma, mean, inv_std = grp[2:]
                    s_ = gamma * inv_std
                    W, b = W * s_.view(-1, 1, 1, 1), b * s_ + (beta - mean * s_)
                flat += [W, b]
            params = flat
        assert len(params) == 4 * self.total, 'expected %d arrays, got %d' % (4 * self.total, len(params))
        # filters per level: n_filters * 2**p, p < 6 (models/fcn_down.py:96-99)
        self.filters = []
        f = n_filters
        for p in range(self.total):
            if p < 6:
                f = n_filters * (2 ** p)
            self.filters.append(f)
        assert all(f % 64 == 0 for f in self.filters), 'n_filters must be a multiple of 64'
        self.down, self.up = [], []
        self.hproj_w = None
        cin_real, cin_pad = n_classes, self.y_cpad
        for p in range(self.total):
            W, b = params[2 * p], params[2 * p + 1]
            if p == self.n_pool:
                # First conv after ConcatLayer((h, pool_n)) (h channels first, models/model_helpers.py:93-94).
                # conv(concat(h, x)) = [conv_h(h) + b] + conv_x(x): the bracket does not change between
                # iterations, so it is computed once per batch in fp32 (`hproj`) and enters the per-iteration
                # conv as an fp32 epilogue addend; the loop only runs the x half of the K range.
                Wt = torch.as_tensor(W)
                self.hproj_w = pack_conv(Wt[:, :self.nb_h], b, [(self.nb_h, self.h_pad)], self.filters[p], self.device,
                                         split=self.split)
                Wx, _ = pack_conv(Wt[:, self.nb_h:], b, [(cin_real, cin_pad)], self.filters[p], self.device,
                                  split=self.split)
                self.down.append((Wx, torch.zeros(self.filters[p], dtype=torch.float32, device=self.device)))
            else:
                self.down.append(pack_conv(W, b, [(cin_real, cin_pad)], self.filters[p], self.device, split=self.split))
            cin_real = cin_pad = self.filters[p]
            if self.cbp > 1:
                f_ = self.filters[p]
                self.down_extra[p] = [pack_conv(extra[p][2 * i], extra[p][2 * i + 1], [(f_, f_)], f_, self.device, split=self.split)
                                      for i in range(self.cbp - 1)]
        up_in = self.filters[-1]
        for i, p in enumerate(range(self.total, 0, -1)):
            W, b = params[2 * (self.total + i)], params[2 * (self.total + i) + 1]
            n_cl = n_classes if p == 1 else self.filters[p - 2]   # models/fcn_up.py:29-34
            cout_pad = 16 if p == 1 else n_cl
            if unpool_type == 'standard':
                # Deconv2DLayer W (in, out, 4, 4), flip_filters=False = conv_transpose2d with the flipped kernel Wf:
                #   out[2m + py] = x[m - 1] * Wf[py + 2] + x[m] * Wf[py]
                # i.e. output phase (py, px) is a 2x2 cross-correlation with pad 1 whose taps are (Wf[q + 2], Wf[q])
                Wf = torch.as_tensor(W).flip(2, 3)
                assert tuple(Wf.shape) == (up_in, n_cl, 4, 4), tuple(Wf.shape)
                phases = []
                for py in range(2):
                    row = []
                    for px in range(2):
                        Wc = Wf[:, :, [py + 2, py]][:, :, :, [px + 2, px]].permute(1, 0, 2, 3).contiguous()      # (out, in, 2, 2)
                        row.append(pack_conv(Wc, b, [(up_in, up_in)], cout_pad, self.device, split=self.split_up))
                    phases.append(row)
                self.up.append(phases)
            else:
                self.up.append(pack_conv(W, b, [(up_in, up_in)], cout_pad, self.device, split=self.split_up))
            up_in = n_cl
        # up_conv1 (64 -> n_classes at full resolution, bf16 operands): the N-packed filter bank (iiseg_conv_desc.weight_npack)
        self.up1_npack = None
        if unpool_type != 'standard' and not self.split_up and self.filters[0] == 64:
            self.up1_npack = pack_npack16(self.up[-1][0])
        # the softmax tail + update can run in the epilogue of the last conv when that conv is a bf16 3x3 conv
        self.fusable_update = unpool_type != 'standard' and not self.split_up
        self._ws = {}

    # -- shapes -----------------------------------------------------------
    def level_sizes(self, H, W):
        """Pre-pool spatial size of every level (conv output), SURVEY.md App. B.1."""
        sizes = []
        pad = self.padding if self.padding > 0 else 1
        h, w = H + 2 * pad - 2, W + 2 * pad - 2
        for p in range(self.total):
            sizes.append((h, w))
            h, w = h // 2, w // 2
        return sizes

    def h_spatial(self, H, W):
        if self.n_pool == 0:
            return H, W
        s = self.level_sizes(H, W)[self.n_pool - 1]
        return s[0] // 2, s[1] // 2

    def cone_windows(self, H, W):
        """Dependency cone of the final centre crop through the expanding path.

        Only the H x W centre of `up_conv1` is kept (CroppingLayer, models/fcn_up.py:106-113), so each
        upper layer is needed only where that crop can see it: conv window Wc[p] at level p needs the
        unpooled map on Wu[p] = Wc[p] dilated by the 3x3 halo (clipped to the map), which needs the level
        below on the pool windows it touches.  With pad=100 this is about half of every map.  Values outside
        the cone never reach the output, so restricting the work is exact.  Returns {p: (h_lo, h_hi, w_lo,
        w_hi)} for Wc and Wu, p = 1..P, in level-p (pre-pool) coordinates."""
        sizes = self.level_sizes(H, W)
        oh, ow = (sizes[0][0] - H) // 2, (sizes[0][1] - W) // 2
        Wc, Wu = {1: (oh, oh + H, ow, ow + W)}, {}
        for p in range(1, self.total + 1):
            hl, hh, wl, wh = Wc[p]
            Sh, Sw = sizes[p - 1]
            Wu[p] = (max(hl - 1, 0), min(hh + 1, Sh), max(wl - 1, 0), min(wh + 1, Sw))
            if p < self.total:
                ul, uh, vl, vh = Wu[p]
                S2h, S2w = sizes[p]
                Wc[p + 1] = (ul // 2, min((uh - 1) // 2 + 1, S2h), vl // 2, min((vh - 1) // 2 + 1, S2w))
        return Wc, Wu

    def down_windows(self, H, W):
        """y-dependent window of every contracting-path conv output, even-aligned for the fused pool.

        A level-p output pixel outside Wu[p] (the same cone, walked downwards from the image) sees only
        the zero padding around y, constant borders of the levels above and -- from level n_pool+1 on -- h,
        none of which change between iterations.  So after the first iteration of a batch, pool_p and
        mask_p outside the window already hold the right values and only the window is recomputed."""
        sizes = self.level_sizes(H, W)
        _, Wu = self.cone_windows(H, W)
        D = {}
        for p in range(1, self.total + 1):
            hl, hh, wl, wh = Wu[p]
            Sh, Sw = sizes[p - 1]
            D[p] = (hl & ~1, min((hh + 1) & ~1, (Sh // 2) * 2), wl & ~1, min((wh + 1) & ~1, (Sw // 2) * 2))
        return D

    def executed_conv_flops(self, H, W, steady_state=True, tensor=False):
        """Executed algorithmic FLOPs (2*MAC, real channel counts) of the 2P conv launches of one
        application, per image: cone windows on the expanding path; on the contracting path the
        y-dependent windows (`steady_state`, iterations 2..N) or the full maps (first iteration).
        `tensor`: the FLOPs the tensor cores execute -- three bf16 products per fp32-accurate MAC on the
        layers that run split precision ('fp32x3', 'mixed')."""
        sizes = self.level_sizes(H, W)
        Wc, _ = self.cone_windows(H, W)
        D = self.down_windows(H, W)
        fl = []
        cin = self.n_classes
        for p in range(self.total):
            hl, hh, wl, wh = D[p + 1] if steady_state else (0, sizes[p][0], 0, sizes[p][1])
            f = 2.0 * (hh - hl) * (wh - wl) * cin * self.filters[p] * 9
            if p == self.n_pool and not steady_state:    # the hoisted h half runs once, with the first iteration
                f += 2.0 * sizes[p][0] * sizes[p][1] * self.nb_h * self.filters[p] * 9
            fl.append(f * (3.0 if (tensor and self.split) else 1.0))
            cin = self.filters[p]
        up_in = self.filters[-1]
        for p in range(self.total, 0, -1):
            n_cl = self.n_classes if p == 1 else self.filters[p - 2]
            hl, hh, wl, wh = Wc[p]
            fl.append(2.0 * (hh - hl) * (wh - wl) * up_in * n_cl * 9 * (3.0 if (tensor and self.split_up) else 1.0))
            up_in = n_cl
        return fl

    def workspace(self, B, H, W):
        """Activation buffers for one application, allocated once per (B, H, W) and
        kept resident (they are baked into the captured CUDA graph)."""
        key = (B, H, W)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev, bf = self.device, torch.bfloat16
        sizes = self.level_sizes(H, W)
        Wc, Wu = self.cone_windows(H, W)
        ws = {'pool': [], 'mask': [], 'unpool': {}, 'upconv': {}, 'Wc': Wc, 'Wu': Wu}
        for p, (h, w) in enumerate(sizes):
            f = self.filters[p]
            ws['pool'].append(torch.empty((B, h // 2, w // 2, self.cm * f), dtype=bf, device=dev))
            ws['mask'].append(torch.empty((B, h // 2, w // 2, f // 8), dtype=torch.int32, device=dev))
        for p in range(1, self.total + 1):
            ul, uh, vl, vh = Wu[p]
            # zeros: with the DePool2D-in-the-producer path the trailing odd row / column of a level is never written
            ws['unpool'][p] = torch.zeros((B, uh - ul, vh - vl, self.cm_up * self.filters[p - 1]), dtype=bf, device=dev)
            if p > 1:   # up_conv_p output: level-p window, channels of level p-1; skip partner pool_{p-1} has size S_p
                hl, hh, wl, wh = Wc[p]
                assert sizes[p - 1] == tuple(ws['pool'][p - 2].shape[1:3]), 'skip-sum needs equal sizes'
                ws['upconv'][p] = torch.empty((B, hh - hl, wh - wl, self.cm_up * self.filters[p - 2]), dtype=bf, device=dev)
        hp = sizes[self.n_pool]
        ws['hproj'] = torch.empty((B, hp[0], hp[1], self.filters[self.n_pool]), dtype=torch.float32, device=dev)
        ws['logits'] = torch.empty((B, H, W, 16), dtype=torch.float32, device=dev)
        self._ws[key] = ws
        return ws

    # -- one application ----------------------------------------------------
    def _bn_mask_level(self, ws, x, p, pools, pad, kw, sizes, per_image, pm):
        """One level of the bn=1 mask pass: conv + rectify to an fp32 map, its per-channel batch statistics
        (iiseg_channel_stats: mean and 1 / sqrt(biased variance + 1e-4) over batch, rows, cols), then the same conv again with
        (x - mean) * (gamma * inv_std) + beta, the 2x2 pool and the tie mask in its epilogue.  `per_image`: statistics per
        image -- the reference's loop calls de_fn on one image at a time (iterative_inference.py:261-268)."""
        B = x.shape[0]
        Wk, bk = self.down[p]
        beta, gamma = self.bn_gb[p]
        hh, ww = sizes[p]
        C_ = self.filters[p]
        key = ('bn_z', p, B)
        if key not in ws:
            need = K._lib.load().iiseg_channel_stats_chunks(B, hh, ww) * C_ * 2
            ws[key] = (torch.empty((B, hh, ww, self.cm * C_), dtype=torch.bfloat16, device=self.device),
                       torch.empty((B, hh, ww, C_), dtype=torch.float32, device=self.device),
                       torch.empty((C_,), dtype=torch.float32, device=self.device),
                       torch.empty((C_,), dtype=torch.float32, device=self.device),
                       torch.empty((need,), dtype=torch.float64, device=self.device))
        zb, z, mean, inv_std, scratch = ws[key]
        for a, b in ([(i, i + 1) for i in range(B)] if per_image else [(0, B)]):
            kwg = dict(kw)
            if 'addend' in kwg:
                kwg['addend'] = kwg['addend'][a:b]
            K.conv2d(x[a:b], Wk, bk, 3, 3, pad, relu=True, out=zb[a:b], split=self.split, **kwg)
            K.widen_nhwc(zb[a:b], z[a:b], split=self.split)        # (the (hi | lo) pair back to) one fp32 value per element
            K.channel_stats(z[a:b], 0, C_, mean, inv_std, scratch, eps=1e-4)
            K.conv2d(x[a:b], Wk, bk, 3, 3, pad, relu=True, pooled=pools[p][a:b], pool_mask=pm[a:b] if pm is not None else None,
                     split=self.split, post_affine=(gamma * inv_std, beta, mean), **kwg)

    def logits(self, h_bf16, y_bf16, full_down=True, update=None, y_f32=None, noise=None, per_image_stats=False):
        """h_bf16: NHWC bf16 (B, Hh, Wh, h_pad); y_bf16: NHWC bf16 (B, H, W, y_cpad).
        Returns fp32 NHWC16 logits of the centre-crop window (B, H, W, 16).
        `full_down=False` recomputes only the y-dependent windows of the contracting path; valid when
        the workspace already holds a full pass for the same h (see `down_windows`).
        `update` (bf16 expanding path: 'bf16' and 'mixed'): dict(y, active, norm_acc, step) -- the softmax tail and the
        iterative-inference update run in the epilogue of the last conv (y and y_bf16 are updated in
        place, the logits are never stored) and None is returned.
        `y_f32` (NCHW fp32, the values y_bf16 was packed from) and optionally `noise` (a list of `total` N(0,1) tensors of that
        shape, one per DePool2D, level 1 first; drawn here when None) are needed when the net was built with mask_noise > 0."""
        B, H, W, _ = y_bf16.shape
        if self.unpool_type == 'standard' or self.cbp > 1 or self.bn_batch_masks:
            full_down = True          # (no y-dependent windows for these variants: every level is computed in full)
        ws = self.workspace(B, H, W)
        sizes = self.level_sizes(H, W)
        Wc, Wu = ws['Wc'], ws['Wu']
        assert tuple(h_bf16.shape) == (B,) + self.h_spatial(H, W) + (self.cm * self.h_pad,), \
            (tuple(h_bf16.shape), self.h_spatial(H, W), self.h_pad)
        assert y_bf16.shape[3] == self.cm * self.y_cpad
        sp, spu = self.split, self.split_up
        mixed = sp and not spu       # pair tensors on the way down, plain bf16 on the way up
        x = y_bf16
        D = None if full_down else self.down_windows(H, W)
        if full_down:       # new h: the iteration-invariant half of the concat conv, once per batch, fp32
            K.conv2d(h_bf16, self.hproj_w[0], self.hproj_w[1], 3, 3, 1, relu=False, out=ws['hproj'], out_f32=True,
                     split=sp)
        # (input, role, levels computed, the one level whose mask this pass writes or None = all)
        passes = [(x, True, self.total, None)]
        if self.mask_noise > 0:
            # DePool2D's mask sub-graphs: every DePool2D calls lasagne.layers.get_output(...) itself (layers/mylayers.py:91-93)
            # and every symbolic call of the GaussianNoiseLayer is an independent random stream, so level p's tie mask comes
            # from ITS OWN pass of the contracting path (levels 1..p) on y + sigma * N_p -- observed by executing the reference
            # with logged draws (tests/golden/ref_noise.npz).  The pooled values of those passes go to a scratch set; the value
            # pass below leaves the masks alone.  `noise`: a list of `total` N(0,1) tensors (level 1 first); None: drawn here;
            # a single tensor: one shared pass for all levels (the same marginal distribution per level, not the same joint).
            assert y_f32 is not None, 'mask_noise > 0 needs the fp32 y (y_f32=) to draw the noised input of the mask pass'
            if noise is None:
                noise = [torch.randn(y_f32.shape, dtype=torch.float32, device=y_f32.device) for _ in range(self.total)]
            if 'pool_m' not in ws:
                ws['pool_m'] = [torch.empty_like(t) for t in ws['pool']]
            if isinstance(noise, (list, tuple)):
                assert len(noise) == self.total
                passes = [(K.noise_pack(y_f32, nz, self.mask_noise, self.y_cpad, split=sp), 'masks', lvl + 1, lvl) for lvl, nz in enumerate(noise)]
            else:
                passes = [(K.noise_pack(y_f32, noise, self.mask_noise, self.y_cpad, split=sp), 'masks', self.total, None)]
            passes.append((x, 'values', self.total, None))
        elif self.bn_batch_masks:
            if 'pool_m' not in ws:
                ws['pool_m'] = [torch.empty_like(t) for t in ws['pool']]
            passes = [(x, 'masks', self.total, None), (x, 'values', self.total, None)]
        for x, role, upto, only in passes:
            pools = ws['pool_m'] if role == 'masks' else ws['pool']
            for p in range(upto):
                Wk, bk = self.down[p]
                pad = self.padding if (p == 0 and self.padding > 0) else 1
                win = None
                if D is not None:
                    hl, hh, wl, wh = D[p + 1]
                    win = (hl, wl, hh - hl, wh - wl)
                # conv + ReLU with Pool2DLayer(2) and the DePool2D tie mask fused in the epilogue: the
                # pre-pool map is consumed on chip and never written (nothing else reads it)
                pm = ws['mask'][p] if (role is True or (role == 'masks' and only in (None, p))) else None
                kw = {}
                if p == self.n_pool:
                    kw = dict(addend=ws['hproj'], addend_off=(win[0], win[1]) if win else (0, 0))
                if role == 'masks' and self.bn_batch_masks:
                    self._bn_mask_level(ws, x, p, pools, pad, kw, sizes, per_image_stats, pm)
                elif self.cbp > 1:      # conv_p_1 .. conv_p_{k-1} write full pre-pool maps, conv_p_k carries the pool
                    pre = ws.setdefault(('pre', p), [None, None])
                    hh_, ww_ = sizes[p]
                    for i in range(self.cbp):
                        last = i == self.cbp - 1
                        Wi, bi = (Wk, bk) if i == 0 else self.down_extra[p][i - 1]
                        if last:
                            K.conv2d(x, Wi, bi, 3, 3, 1, relu=True, pooled=pools[p], pool_mask=pm, split=sp)
                        else:
                            if pre[i & 1] is None:
                                pre[i & 1] = torch.empty((B, hh_, ww_, self.cm * self.filters[p]), dtype=torch.bfloat16, device=self.device)
                            x = K.conv2d(x, Wi, bi, 3, 3, pad if i == 0 else 1, relu=True, out=pre[i & 1], split=sp, **(kw if i == 0 else {}))
                else:
                    K.conv2d(x, Wk, bk, 3, 3, pad, relu=True, window=win, pooled=pools[p], pool_mask=pm, split=sp,
                             post_affine=self.post[p], **kw)
                x = pools[p]
        if self.unpool_type == 'standard':
            return self._up_standard(ws, sizes, B, H, W)
        u, u_origin = ws['pool'][-1], (0, 0)
        prefilled = False
        for i, p in enumerate(range(self.total, 0, -1)):
            h, w = sizes[p - 1]
            ul, uh, vl, vh = Wu[p]
            hl, hh, wl, wh = Wc[p]
            Wk, bk = self.up[i]
            if p == 1 and self.fuse_depool and not sp:
                # DePool2D expanded inside the conv's loader: the 64-channel full-resolution map (221 MB per batch of
                # 10) is neither written nor read; the conv runs on the virtual map in full-map coordinates
                dp = dict(window=(hl, wl, hh - hl, wh - wl), depool=(ws['mask'][0], h, w, u_origin))
                if update is not None:
                    K.conv2d(u, Wk, bk, 3, 3, 1, relu=False, out_f32=True,
                             update=dict(update, y_bf16=y_bf16, C=self.n_classes), **dp)
                    return None
                K.conv2d(u, Wk, bk, 3, 3, 1, relu=False, out=ws['logits'], out_f32=True, **dp)
                break
            if prefilled:        # the previous conv's epilogue already wrote DePool2D(u) into this level's window
                up = ws['unpool'][p]
            else:
                # mixed: the first unpool reads the hi halves of the contracting path's pool pair
                up = K.unpool2(u, ws['mask'][p - 1], h, w, out=ws['unpool'][p], u_origin=u_origin,
                               window=(ul, vl, uh - ul, vh - vl), split=(2 if (mixed and i == 0) else spu))
            win = (hl - ul, wl - vl, hh - hl, wh - wl)     # conv window inside the unpooled window tensor
            if p > 1 and self.fuse_depool_out and self.skip and not sp and not (p == 2 and self.fuse_depool):
                # skip-sum, then DePool2D for the next level written straight from this conv's epilogue (u itself is
                # never stored): one launch and one round trip of u through HBM less per level
                nul, _, nvl, _ = Wu[p - 1]
                K.conv2d(up, Wk, bk, 3, 3, 1, relu=False, window=win, addend=ws['pool'][p - 2], addend_off=(hl, wl),
                         depool_out=(ws['unpool'][p - 1], ws['mask'][p - 2], (nul, nvl), (hl, wl)))
                prefilled = True
            elif p > 1:   # skip-sum with pool_{p-1} (full map) read at the window offset
                prefilled = False
                u = K.conv2d(up, Wk, bk, 3, 3, 1, relu=False, window=win, addend=ws['pool'][p - 2] if self.skip else None,
                             addend_off=(hl, wl), out=ws['upconv'][p], split=spu, addend_pair_hi=mixed and self.skip)
                u_origin = (hl, wl)
            else:       # centre crop (CroppingLayer, layers/mylayers.py:36-57): exactly the H x W window
                if update is not None:
                    K.conv2d(up, Wk, bk, 3, 3, 1, relu=False, window=win, out_f32=True, weight_npack=self.up1_npack,
                             update=dict(update, y_bf16=y_bf16, C=self.n_classes, y_split=sp))
                    return None
                K.conv2d(up, Wk, bk, 3, 3, 1, relu=False, window=win, out=ws['logits'], out_f32=True, split=spu,
                         weight_npack=self.up1_npack)
        return ws['logits']


def _phase_range(size_in, size_out, crop, q):
    """Phase q of a stride-2, 4-tap transposed conv of a length-`size_in` axis produces out[2m + q], m = 0..size_in; after
    a centre crop by `crop` to `size_out` the kept m are [lo, hi] and out index 2*lo + q - crop is the first written."""
    lo = max(0, (crop - q + 1) // 2)
    hi = min(size_in, (size_out - 1 + crop - q) // 2)
    return lo, hi - lo + 1, 2 * lo + q - crop


def _up_standard(self, ws, sizes, B, H, W):
    """Expanding path of unpool_type='standard' (models/fcn_up.py:37-63): per level Deconv2DLayer(4, stride 2, 'valid',
    linear) as four 2x2 phase convolutions whose epilogues store with pixel stride 2 straight into the centre-cropped
    map and add the skip partner pool_{p-1} read at the same stride (ElemwiseSumLayer, cropping=center)."""
    spu, mixed = self.split_up, self.split and not self.split_up
    prev, prev_pair_hi = ws['pool'][-1], mixed
    for i, p in enumerate(range(self.total, 0, -1)):
        hp, wp = prev.shape[1], prev.shape[2]
        dh, dw = 2 * hp + 2, 2 * wp + 2
        Sh, Sw = sizes[p - 1] if p > 1 else (H, W)
        assert dh >= Sh and dw >= Sw
        ch, cw = (dh - Sh) // 2, (dw - Sw) // 2
        key = ('std', p)
        dest = ws.get(key)
        if dest is None:
            c = 16 if p == 1 else self.cm_up * self.filters[p - 2]
            dest = ws[key] = torch.empty((B, Sh, Sw, c), dtype=torch.float32 if p == 1 else torch.bfloat16, device=self.device)
        for py in range(2):
            m0, mh, o_h0 = _phase_range(hp, Sh, ch, py)
            for px in range(2):
                n0, nw, o_w0 = _phase_range(wp, Sw, cw, px)
                Wk, bk = self.up[i][py][px]
                kw = {}
                if p > 1 and self.skip:       # skip partner, same size as the cropped map (pool_{p-1} never is the larger one)
                    kw = dict(addend=ws['pool'][p - 2], addend_off=(o_h0, o_w0), addend_pair_hi=mixed)
                K.conv2d(prev, Wk, bk, 2, 2, 1, relu=False, window=(m0, n0, mh, nw), out_f32=(p == 1), split=spu,
                         out_strided=(dest, 2, (o_h0, o_w0)), src_pair_hi=prev_pair_hi, **kw)
        prev, prev_pair_hi = dest, False
    return prev


DAENet._up_standard = _up_standard


def buildDAE(input_concat_h_vars, input_mask_var, n_classes, nb_features_to_concat,
             padding, ae_h=False, void_labels=[], path_weights='/Tmp/romerosa/itinf/models/',
             model_name='dae_model.npz', trainable=False, load_weights=False,
             out_nonlin=None, concat_h=['input'], noise=0.1, n_filters=64,
             conv_before_pool=1, additional_pool=0, dropout=0., skip=False,
             unpool_type='standard', bn=0, params=None, precision='bf16', stochastic_masks=None):
    """Same arguments as the reference builder (models/DAE_h.py:12-17); returns the handle
    of 'probs_dimshuffle'.  The symbolic inputs are ignored.  Built for the benchmark
    configuration; other variants raise.  `noise` only matters for training and for the
    reference's non-deterministic mask sub-graph (layers/mylayers.py:91-93); inference here
    is the deterministic noise=0 graph."""
    import os
    import warnings
    if unpool_type not in ('trackind', 'inverse', 'standard'):
        raise ValueError('Unkown unpool type')                       # models/fcn_up.py:115
    if conv_before_pool > 1 and bn:
        raise NotImplementedError('B200 DAE_h: conv_before_pool > 1 is built for bn=0')
    # ae_h (models/DAE_h.py:12-49): names the layers 'h_to_recon' / 'h_hat' and freezes the pre-h parameters for the
    # training loss of train_dae.py:317-320; the graph that produces 'probs_dimshuffle' is unchanged -> nothing to do here.
    # dropout: DropoutLayer is the identity under deterministic=True (iterative_inference.py:189-190), so it does not
    # change inference.  NB the reference builds DePool2D's mask sub-graph WITHOUT deterministic (layers/mylayers.py:91-93):
    # with noise > 0 or dropout > 0 its masks come from a separately noised / dropped-out pass even at test time.  This
    # build is the deterministic graph; warn so that a caller comparing against such a reference run knows.
    if stochastic_masks is None:          # as the reference: its DePool2D mask sub-graphs are noised whenever the DAE has noise > 0
        stochastic_masks = noise > 0
    if unpool_type == 'trackind' and (dropout > 0 or (noise > 0 and not stochastic_masks)):
        warnings.warn('buildDAE: noise=%s dropout=%s only affect the reference\'s non-deterministic DePool2D mask sub-graph at '
                      'inference (layers/mylayers.py:91-93); this build uses the deterministic masks (noise > 0 with stochastic_masks=False was requested, or dropout, whose mask stream is not built; bn=1: the batch-statistics mask pass is always built)' % (noise, dropout), stacklevel=2)
    concat_h = list(concat_h)
    if len(concat_h) != 1 or not (concat_h[-1] == 'input' or concat_h[-1] in ('pool1', 'pool2', 'pool3', 'pool4', 'pool5')):
        # (several entries share ONE nb_features_to_concat in the reference, models/model_helpers.py:91, so they only build
        # there when all conditioning tensors have the same channel count)
        raise NotImplementedError('B200 DAE_h concatenates one conditioning tensor: concat_h=[\'input\'] or [\'poolN\']')
    if params is None:
        if not load_weights:
            raise ValueError('buildDAE needs weights: pass params= or load_weights=True with path_weights')
        params = load_npz_params(os.path.join(path_weights, model_name))
    net = DAENet(n_classes, nb_features_to_concat, padding, params, concat_h=tuple(concat_h),
                 n_filters=n_filters, additional_pool=additional_pool, precision=precision, unpool_type=unpool_type, bn=bn,
                 mask_noise=noise if stochastic_masks else 0.0, skip=skip, conv_before_pool=conv_before_pool)
    return LayerHandle(net, 'probs_dimshuffle', n_classes)

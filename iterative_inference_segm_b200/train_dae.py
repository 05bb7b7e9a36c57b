"""The DAE training step of train_dae.py on the sm_100a kernels.

`DAETrainer.step(h, y, target, noise_main, noise_mask)` is one `train_fn(*H_pred, Y_pred, L)` call
(train_dae.py:334-335, 378): GaussianNoiseLayer -> DAE forward (no `deterministic`) -> masked
crossentropy + lmb * masked squared_error (train_dae.py:279-293, metrics.py:68-91,144-156) ->
gradients of all 24 parameter arrays -> lasagne.updates.rmsprop (train_dae.py:326-327).  Benchmark
configuration: kind='standard', unpool_type='trackind', skip=True, conv_before_pool=1, bn=0, dropout=0.

Where the arithmetic runs:
  * forward: the inference kernels on full maps (weights change every step, so none of the
    iteration hoists of the inference loop apply); with noise > 0 the DePool2D masks come from a
    SECOND stochastically noised pass of the contracting path, as in the reference
    (layers/mylayers.py:91-93 calls get_output without `deterministic`);
  * data gradients: the same tcgen05 conv kernel on the flipped / transposed filter bank;
  * weight (+ bias) gradients: one K-major GEMM per layer on the same kernel (a 1x1 conv whose
    channel axis is the pixel index): dW[co][tap][ci] = sum_p g[p][co] x[p + tap][ci], operands
    written by `iiseg_transpose_shift`, a row of ones giving the bias gradient as one more column;
  * pool / unpool / rectify / loss / rmsprop: streaming kernels of csrc/train.cu.
Gradients travel as bf16 NHWC tensors (fp32 accumulation everywhere), master weights and rmsprop
state are fp32 in the GEMM layout.  Data parallelism (`world` = sharding.World): the loss is a masked mean
over the GLOBAL batch, so ranks all-reduce the two loss denominators between the loss pass and the
gradient pass, then sum their weight-gradient matrices (NCCL) before the update: the step equals the
single-device step on the concatenated batch up to fp32 summation order.
"""
import torch

from . import _kernels as K
from ._packing import _as_f32
from .models.DAE_h import DAENet


def _r64(c):
    return (c + 63) // 64 * 64


class _Layer(object):
    """One conv of the DAE: fp32 master bank [Cout_pad][9][Cin_pad] + rmsprop state + bf16 banks."""

    def __init__(self, W, b, splits, cout_pad, dev, dgrad_range):
        W, b = _as_f32(W, dev), _as_f32(b, dev)
        Cout, Cin = W.shape[0], W.shape[1]
        assert Cin == sum(r for r, _ in splits)
        parts, c0 = [], 0
        for real, padded in splits:
            blk = torch.zeros((cout_pad, 3, 3, padded), dtype=torch.float32, device=dev)
            blk[:Cout, :, :, :real] = W[:, c0:c0 + real].permute(0, 2, 3, 1)
            parts.append(blk)
            c0 += real
        self.cout, self.cout_pad = Cout, cout_pad
        self.cin_pad = sum(pd for _, pd in splits)
        self.w = torch.cat(parts, dim=3).reshape(cout_pad, 9, self.cin_pad).contiguous()
        self.b = torch.zeros((cout_pad,), dtype=torch.float32, device=dev)
        self.b[:Cout] = b
        self.acc, self.acc_b = torch.zeros_like(self.w), torch.zeros_like(self.b)
        self.wb = self.w.reshape(cout_pad, -1).to(torch.bfloat16).contiguous()                 # forward bank
        # weight-gradient GEMM columns: filter row r at stride g_rstride, then (s, ci); bias column; pad to 64.
        # A 16-channel input (the first layer) packs its three column shifts into ONE 64-row group per filter row
        # (3 N tiles of 64 instead of 9 of 16: the g^T operand is streamed 3 times, not 9).
        self.g_rstride = 64 if self.cin_pad == 16 else 3 * self.cin_pad
        self.nb = _r64(3 * self.g_rstride + 1)
        self.bias_col = 3 * self.g_rstride
        # data-gradient bank over the input channels [ci0, ci0 + ci_t): wt[ci][8 - tap][co]
        self.ci0, self.ci_t = dgrad_range if dgrad_range is not None else (0, 0)
        self.wt = None
        if dgrad_range is not None:
            self.co_pad_t = cout_pad if cout_pad == 16 else _r64(cout_pad)
            sub = self.w[:, :, self.ci0:self.ci0 + self.ci_t]                                   # [co, tap, ci]
            wt = torch.zeros((self.ci_t, 9, self.co_pad_t), dtype=torch.float32, device=dev)
            wt[:, :, :cout_pad] = sub.flip(1).permute(2, 1, 0)
            self.wt = wt.reshape(self.ci_t, -1).to(torch.bfloat16).contiguous()
        self.zero_bias_t = torch.zeros((max(self.ci_t, 16),), dtype=torch.float32, device=dev)
        self.grad = None            # fp32 [Cout_pad][nb], the weight-gradient GEMM output of the last step

    def lasagne_arrays(self, splits_real):
        """(W (Cout, Cin, 3, 3), b (Cout,)) in the reference's checkpoint layout."""
        w = self.w.reshape(self.cout_pad, 3, 3, self.cin_pad)
        parts, c0 = [], 0
        for real, padded in splits_real:
            parts.append(w[:self.cout, :, :, c0:c0 + real])
            c0 += padded
        return torch.cat(parts, dim=3).permute(0, 3, 1, 2).contiguous(), self.b[:self.cout].clone()


class DAETrainer(object):
    # SMs left free for NCCL's CTAs during the data-parallel backward (iiseg_reserve_sms).  Measured on 2 x B200
    # (tools/train_dp_overlap.py): no exchange 4.91 ms, one blocking all-reduce 5.36, buckets under backward 5.17 with 0 SMs
    # reserved, 5.18 with 4, 5.34 with 8, 5.35 with 16 -- the lost compute costs more than the overlap gains, so: 0.
    DP_RESERVED_SMS = int(__import__('os').environ.get('IISEG_DP_RESERVED_SMS', '0'))
    BUCKET_BYTES = 32 << 20        # gradient all-reduce bucket size (NVSwitch: sized for launch latency / overlap, not link count)

    def __init__(self, n_classes, nb_features_to_concat, padding, params, concat_h=('pool4',), n_filters=64,
                 additional_pool=2, learning_rate=1e-3, noise=0.5, lmb=1.0, rho=0.9, epsilon=1e-6, device='cuda',
                 optimizer='rmsprop', beta1=0.9, beta2=0.999, adam_epsilon=1e-8, training_loss=('crossentropy', 'squared_error'), ae_h=False):
        K.require_device()
        self.dev = dev = torch.device(device)
        self.C, self.lr, self.sigma, self.lmb, self.rho, self.eps = n_classes, learning_rate, noise, lmb, rho, epsilon
        # lasagne.updates.rmsprop(loss, params, learning_rate=lr) or lasagne.updates.adam(...) with lasagne's defaults
        # (train_dae.py:326-331): adam keeps a first-moment bank per layer and the step counter t / step size a_t on the device
        assert optimizer in ('rmsprop', 'adam'), optimizer
        self.optimizer, self.beta1, self.beta2, self.adam_eps = optimizer, beta1, beta2, adam_epsilon
        self.adam_state = torch.zeros((2,), dtype=torch.float32, device=dev)
        # the loss terms train_dae.py:278-294 adds up: crossentropy, dice (metrics.py:93-113, channel 1), lmb * squared_error
        unknown = sorted(set(training_loss) - set(K.LOSS_TERMS))
        if unknown or not training_loss:
            raise NotImplementedError('B200 train step: training_loss terms %s (built: %s)' % (unknown, sorted(K.LOSS_TERMS)))
        self.terms = sum(K.LOSS_TERMS[t] for t in set(training_loss))
        # ae_h (train_dae.py:238-239,317-319): + squared_error(h_to_recon, h_hat).mean() = the mean square of up_conv_{n_pool+1}
        # over its WHOLE map (h_hat = up_conv + h): the crop cone is widened to the full maps from that level down
        self.ae_h = bool(ae_h)
        if self.ae_h and additional_pool < 1:
            raise ValueError('ae_h needs additional_pool >= 1: no layer is named h_hat otherwise (models/fcn_up.py:145-146)')
        geo = self.geo = DAENet.__new__(DAENet)          # geometry helpers only (level sizes, crop cone)
        geo.n_classes, geo.nb_h, geo.h_pad, geo.padding = n_classes, nb_features_to_concat, _r64(nb_features_to_concat), padding
        last = concat_h[-1]
        geo.n_pool = int(last[-1])
        geo.total = geo.n_pool + additional_pool
        geo.filters = [n_filters * 2 ** min(p, 5) for p in range(geo.total)]
        geo.split, geo.cm, geo.y_cpad = False, 1, 16
        P, f = geo.total, geo.filters
        assert len(params) == 4 * P
        self.down, self.up, self._splits = [], [], []
        cin = n_classes
        for p in range(P):
            if p == geo.n_pool:
                splits = [(geo.nb_h, geo.h_pad), (cin, cin)]
                rng = (geo.h_pad, cin)                     # the conv propagates to its own pool half only
            else:
                splits = [(cin, 16 if p == 0 else cin)]
                rng = (0, cin) if p > 0 else None          # no gradient w.r.t. the network input
            self.down.append(_Layer(params[2 * p], params[2 * p + 1], splits, f[p], dev, rng))
            self._splits.append(splits)
            cin = f[p]
        up_in = f[-1]
        for i, p in enumerate(range(P, 0, -1)):
            n_cl = n_classes if p == 1 else f[p - 2]
            self.up.append(_Layer(params[2 * (P + i)], params[2 * (P + i) + 1], [(up_in, up_in)], 16 if p == 1 else n_cl, dev,
                                  (0, up_in)))
            self._splits.append([(up_in, up_in)])
            up_in = n_cl
        self.sums = torch.zeros((10,), dtype=torch.float64, device=dev)      # 0-3 CE / MSE, 4-6 dice, 8-9 ae_h
        self._graphs = {}
        self.last_loss = None
        # One flat fp32 buffer holds the 24 weight-gradient matrices in the order backward produces them (expanding path
        # from the output, then the contracting path from the bottleneck): the data-parallel step all-reduces it in a
        # few large buckets, each launched as soon as its last layer's gradient is written, under the rest of backward.
        order = [self.up[P - p] for p in range(1, P + 1)] + [self.down[p - 1] for p in range(P, 0, -1)]
        self._grad_flat = torch.zeros((sum(l.cout_pad * l.nb for l in order),), dtype=torch.float32, device=dev)
        off = 0
        for l in order:
            l.grad = self._grad_flat[off:off + l.cout_pad * l.nb].view(l.cout_pad, l.nb)
            l.grad_off = off
            off += l.cout_pad * l.nb
        self._buckets, start, last = [], 0, None            # [(lo, hi, id of the layer that completes the bucket)]
        for l in order:
            end = l.grad_off + l.cout_pad * l.nb
            if (end - start) * 4 >= self.BUCKET_BYTES or l is order[-1]:
                self._buckets.append((start, end, id(l)))
                start = end
        self._works = []

    def layers(self):
        return self.down + self.up

    def params(self):
        """The 24 arrays in the reference's checkpoint order (np.savez(*get_all_param_values), train_dae.py:436-445)."""
        out = []
        for lay, splits in zip(self.layers(), self._splits):
            out += list(lay.lasagne_arrays(splits))
        return out

    # ------------------------------------------------------------------ forward
    border_once = True        # levels above h: the y-independent border is computed on one image and copied (see _down_level)

    def _down_level(self, p, lay, x, h):
        """Contracting level p (0-based) for the batch x: conv3x3 + ReLU with the 2x2 pool, the tie mask and the exact-zero mask
        fused in the epilogue.  Above the h concat, a pixel outside the y-dependent window (DAENet.down_windows: the image, dilated
        by one pixel per conv, inside the pad-100 border) sees only the zero padding around y and the constant borders of the
        levels above: it is the same for EVERY image of the batch.  So the full map is computed for image 0 only and copied
        (data movement), and the other images run the window -- ~30 % of the map with padding 100.  Bit-identical to full
        launches: a windowed launch produces the values the full launch produces there, and the border values do not depend on
        the image (tests/test_train_gpu.py::test_border_once_equals_full_launches)."""
        geo, sizes = self.geo, self._sizes
        n = x.shape[0]
        hh, ww = sizes[p]
        pooled = torch.empty((n, hh // 2, ww // 2, lay.cout_pad), dtype=torch.bfloat16, device=self.dev)
        mask = torch.empty((n, hh // 2, ww // 2, lay.cout_pad // 8), dtype=torch.int32, device=self.dev)
        zmask = torch.empty_like(mask)        # exact zeros before the rectifier (gradient 0.5 there)
        pad = geo.padding if (p == 0 and geo.padding > 0) else 1
        if p == geo.n_pool:
            K.conv2d(h, lay.wb, lay.b, 3, 3, pad, relu=True, src1=x, pooled=pooled, pool_mask=mask, pool_zmask=zmask)
            return pooled, mask, zmask
        win = self._dwin.get(p) if self.border_once else None
        if win is None or n == 1:
            K.conv2d(x, lay.wb, lay.b, 3, 3, pad, relu=True, pooled=pooled, pool_mask=mask, pool_zmask=zmask)
        else:
            K.conv2d(x[:1], lay.wb, lay.b, 3, 3, pad, relu=True, pooled=pooled[:1], pool_mask=mask[:1], pool_zmask=zmask[:1])
            for t in (pooled, mask, zmask):
                K.broadcast_image(t)
            K.conv2d(x[1:], lay.wb, lay.b, 3, 3, pad, relu=True, window=win, pooled=pooled[1:], pool_mask=mask[1:], pool_zmask=zmask[1:])
        return pooled, mask, zmask

    def _down_windows(self, H, W):
        """{p (0-based): (oh0, ow0, OH, OW)} for the levels above the h concat whose y-dependent window is worth a separate launch."""
        geo, sizes = self.geo, self._sizes
        D, out = geo.down_windows(H, W), {}
        for p in range(geo.n_pool):
            hl, hh, wl, wh = D[p + 1]
            if (hh - hl) * (wh - wl) <= 0.7 * sizes[p][0] * sizes[p][1]:
                out[p] = (hl, wl, hh - hl, wh - wl)
        return out

    def _down(self, x0, h, upto=None):
        pools, masks, zmasks = [], [], []
        x = x0
        for p, lay in enumerate(self.down[:upto]):
            pooled, mask, zmask = self._down_level(p, lay, x, h)
            pools.append(pooled); masks.append(mask); zmasks.append(zmask)
            x = pooled
        return pools, masks, zmasks

    batched_mask_passes = True

    def _down_merged(self, y, noise_main, noise_mask, h):
        """The main pass and the P noised mask passes as ONE batch per level: [pass l | passes l+1..P | main pass] at level l, so
        that a level is one launch (plus the shared border, `_down_level`) instead of two; the first B images leave after their
        level's mask is taken, the main pass rides at the end through all levels.  Per-image results do not depend on the batch,
        so everything is bit-identical to separate passes.  Returns x0, pools, masksA, zmasks (main pass) and masksB."""
        geo = self.geo
        P, B = geo.total, y.shape[0]
        x = torch.empty(((P + 1) * B, y.shape[2], y.shape[3], 16), dtype=torch.bfloat16, device=self.dev)
        for lvl in range(P):          # y + sigma * N_l straight into pass l's slot (no P-fold copy of y)
            K.noise_pack(y, noise_mask[lvl], self.sigma, 16, out=x[lvl * B:(lvl + 1) * B])
        K.noise_pack(y, noise_main, self.sigma, 16, out=x[P * B:])
        x0 = x[P * B:]
        pools, masksA, zmasks, masksB = [], [], [], []
        for p, lay in enumerate(self.down):
            pooled, mask, zmask = self._down_level(p, lay, x, h.repeat(P - p + 1, 1, 1, 1) if p == geo.n_pool else None)
            masksB.append(mask[:B])
            pools.append(pooled[-B:]); masksA.append(mask[-B:]); zmasks.append(zmask[-B:])
            x = pooled[B:]
        return x0, pools, masksA, zmasks, masksB

    def forward(self, h_bf16, y, noise_main=None, noise_mask=None, forced=None):
        """Training-mode forward; keeps what the backward pass needs.  Returns fp32 NHWC16 logits.
        `noise_mask`: [P, B, C, H, W] = one N(0,1) tensor per DePool2D, level 1 first (the reference's graph); [B, C, H, W] = one
        shared mask pass; None = masks from the main pass.
        `forced` (parity tooling): dict(masksA, zmasks, masksB[, positive]) of per-level tensors that REPLACE the discrete
        decisions this pass takes -- which window elements are maxima (tie masks, nibble layout), which pre-rectifier
        values are exactly zero, and (`positive`: bool [B,h/2,w/2,C] per level) whether a window's maximum is positive, i.e.
        whether the rectifier passes the gradient -- so that only the arithmetic is this path's.  The pooled values of
        windows whose forced sign differs (all within rounding of zero) are nudged to the forced side of zero."""
        geo = self.geo
        B, _, H, W = y.shape
        self._sizes = sizes = geo.level_sizes(H, W)
        self.Wc, self.Wu = geo.cone_windows(H, W)
        self._dwin = self._down_windows(H, W)
        if self.ae_h:
            for p in range(geo.n_pool + 1, geo.total + 1):
                self.Wc[p] = self.Wu[p] = (0, sizes[p - 1][0], 0, sizes[p - 1][1])
        st = self.st = {'B': B, 'H': H, 'W': W, 'h': h_bf16}
        per_depool = noise_mask is not None and noise_mask.dim() == 5
        if per_depool:
            # the DePool2D mask sub-graphs as the reference's graph has them: every DePool2D re-evaluates the contracting path
            # up to its own pool with an independent noise draw (layers/mylayers.py:91-93; tests/golden/ref_noise.npz), so
            # level p's mask comes from a pass over levels 1..p on y + sigma * noise_mask[p - 1]
            assert noise_mask.shape[0] == geo.total
        if per_depool and self.batched_mask_passes:
            st['x0'], st['pools'], st['masksA'], st['zmasks'], st['masksB'] = self._down_merged(y, noise_main, noise_mask, h_bf16)
        else:
            st['x0'] = K.noise_pack(y, noise_main, self.sigma, 16)
            st['pools'], st['masksA'], st['zmasks'] = self._down(st['x0'], h_bf16)
            if per_depool:               # one launch sequence per pass (the unbatched form of _down_merged, kept for tests)
                st['masksB'] = [self._down(K.noise_pack(y, noise_mask[lvl], self.sigma, 16), h_bf16, upto=lvl + 1)[1][lvl]
                                for lvl in range(geo.total)]
            elif noise_mask is not None:     # one shared, separately noised contracting path for all levels
                _, st['masksB'], _ = self._down(K.noise_pack(y, noise_mask, self.sigma, 16), h_bf16)
            else:
                st['masksB'] = st['masksA']
        if forced is not None:
            if st['masksB'] is st['masksA']:
                st['masksB'] = [m.clone() for m in st['masksA']]
            for key, mine in (('masksA', st['masksA']), ('zmasks', st['zmasks']), ('masksB', st['masksB'])):
                for m, f in zip(mine, forced[key]):
                    m.copy_(f)
            for pooled, pos in zip(st['pools'], forced.get('positive', [])):
                tiny = torch.full_like(pooled, 2.0 ** -100)
                pooled.copy_(torch.where(pos, torch.maximum(pooled, tiny), torch.zeros_like(pooled)))
        P = geo.total
        st['v'] = {}
        u, u_origin = st['pools'][-1], (0, 0)
        for i, p in enumerate(range(P, 0, -1)):
            hh, ww = sizes[p - 1]
            ul, uh, vl, vh = self.Wu[p]
            hl, hh2, wl, wh = self.Wc[p]
            v = K.unpool2(u, st['masksB'][p - 1], hh, ww, u_origin=u_origin, window=(ul, vl, uh - ul, vh - vl))
            st['v'][p] = v
            lay = self.up[i]
            win = (hl - ul, wl - vl, hh2 - hl, wh - wl)
            if self.ae_h and p == geo.n_pool + 1:
                # h_hat = up_conv_p + pool_{p-1} over the full map; the conv's own output is kept for the ae_h term
                assert lay.cout == lay.cout_pad and (hl, wl) == (0, 0)
                st['ae_c'] = K.conv2d(v, lay.wb, lay.b, 3, 3, 1, relu=False, window=win)
                u = K.add_bf16(st['ae_c'], st['pools'][p - 2])
                u_origin = (0, 0)
            elif p > 1:
                u = K.conv2d(v, lay.wb, lay.b, 3, 3, 1, relu=False, window=win, addend=st['pools'][p - 2], addend_off=(hl, wl))
                u_origin = (hl, wl)
            else:
                st['logits'] = K.conv2d(v, lay.wb, lay.b, 3, 3, 1, relu=False, window=win, out_f32=True)
        return st['logits']

    # ------------------------------------------------------------------ backward
    def _wgrad(self, lay, g, x_srcs, g_origin_in_x, pad, bias_from=None):
        """dW (+ db) of `lay` = one GEMM.  g: [B,GH,GW,Cg] gradient w.r.t. the conv output over a window whose origin
        sits at `g_origin_in_x` in the coordinates of the input tensors; x_srcs: [(tensor, real_channels_padded)].
        `bias_from`: the gradient tensor the bias sum runs over when g is only the part that meets a non-zero input."""
        B, GH, GW, Cg = g.shape
        # Both operands live on ONE zero-padded pixel grid of Gh x Gw per image (Gw a multiple of 8): g at the origin, x
        # shifted by `pad`, so that output pixel k meets tap (r, s) at column k + r*Gw + s of x^T.  TMA wants 16-byte
        # aligned K coordinates, so x^T is written three times (one copy per horizontal shift s, a flat shift of the
        # padded grid: where it wraps a row, g is zero) and tap (r, s) is copy s
        # read at K + r*Gw (iiseg_conv_desc.w_groups): 3 transposed copies instead of 9.
        Gh, Gw = GH + 2, (GW + 2 + 7) // 8 * 8
        Pn = B * Gh * Gw
        # split K (the pixel axis) so that the GEMM has a few hundred tiles: the high-resolution layers have tiny M x N
        bn = 64 if lay.cin_pad == 16 else min(lay.cin_pad, 256)
        tiles = max(1, (Cg + 127) // 128) * (3 * lay.g_rstride // bn)
        slabs = max(1, min(64, 296 // tiles, Pn // 4096))
        ldo = (Pn + 64 * slabs - 1) // (64 * slabs) * (64 * slabs)
        gT = torch.empty((Cg, ldo), dtype=torch.bfloat16, device=self.dev)
        K.transpose_shift(g, Cg, (0, 0), (Gh, Gw), (0, 0), gT, 0)
        merged = lay.g_rstride != 3 * lay.cin_pad
        xT = (torch.zeros if merged else torch.empty)((lay.g_rstride, ldo), dtype=torch.bfloat16, device=self.dev)
        row = 0
        for x, c in x_srcs:        # one read of x, three writes: copy s is the same matrix one pixel further along the grid
            K.transpose_shift(x, c, g_origin_in_x, (Gh, Gw), (-pad, -pad), xT, row, nshift=3, shift_rows=lay.cin_pad)
            row += c
        assert row == lay.cin_pad
        if merged:
            K.wgrad_gemm(gT, xT, lay.g_rstride, [(0, r * Gw) for r in range(3)], slabs, lay.nb, out=lay.grad)
        else:
            groups = [(s_ * lay.cin_pad, r * Gw) for r in range(3) for s_ in range(3)]
            K.wgrad_gemm(gT, xT, lay.cin_pad, groups, slabs, lay.nb, out=lay.grad)
        K.bias_grad(g if bias_from is None else bias_from, lay.grad, lay.bias_col)
        if self._dp_world is not None:           # data parallel: this layer may complete a bucket -> all-reduce it now
            for lo, hi, last in self._buckets:
                if last == id(lay):
                    self._works.append(self._dp_world.allreduce_sum_async(self._grad_flat[lo:hi]))
        return lay.grad

    def _dgrad(self, lay, g, window, addend=None):
        """Gradient w.r.t. the conv input over `window` = (j0h, j0w, OH, OW) of the pad-2 correlation of g with the
        flipped / transposed bank (position q relative to g's origin is output index q + 1)."""
        return K.conv2d(g, lay.wt, lay.zero_bias_t[:lay.ci_t].contiguous(), 3, 3, 2, relu=False, window=window, addend=addend)

    _dp_world = None
    keep_grads = None

    def backward(self, target, world=None, overlap=False):
        """`overlap` (data parallel): gradient buckets are all-reduced asynchronously as backward completes them; the
        caller waits on `self._works` before the update."""
        self._dp_world = world if (world is not None and overlap) else None
        self._works = []
        st, geo = self.st, self.geo
        B, H, W = st['B'], st['H'], st['W']
        sizes, Wc, Wu, P = self._sizes, self.Wc, self.Wu, geo.total
        if self.ae_h:
            self._ae_sums()
        if world is None:
            g_c = K.loss_grad(st['logits'], target, self.C, self.lmb, self.sums, terms=self.terms)      # dL/dlogits over Wc[1]
        else:       # the loss is a masked mean over the GLOBAL batch (metrics.py:88-89,153-154): global denominators (and dice sums)
            g_c = K.loss_grad(st['logits'], target, self.C, self.lmb, self.sums, passes=1, terms=self.terms)
            world.allreduce_sum(self.sums)
            g_c = K.loss_grad(st['logits'], target, self.C, self.lmb, self.sums, dlogits=g_c, passes=2, terms=self.terms)
        skip = {}                                                                      # level p -> (grad of pool_p from the skip-sum, its window)
        g_u = None
        for i, p in enumerate(range(1, P + 1)):          # expanding path, output towards the bottleneck
            lay = self.up[P - p]
            hl, hh, wl, wh = Wc[p]
            ul, uh, vl, vh = Wu[p]
            v = st['v'][p]
            self._wgrad(lay, g_c, [(v, lay.cin_pad)], (hl - ul, wl - vl), 1)
            g_v = self._dgrad(lay, g_c, (ul - hl + 1, vl - wl + 1, uh - ul, vh - vl))
            # pooled positions under the unpooled window
            S2h, S2w = sizes[p - 1][0] // 2, sizes[p - 1][1] // 2
            pu = (ul // 2, min((uh - 1) // 2 + 1, S2h), vl // 2, min((vh - 1) // 2 + 1, S2w))
            ae_level = self.ae_h and p == geo.n_pool          # the level whose fused sum is h_hat = c_{p+1} + pool_p
            if p < P and not ae_level:
                assert pu == tuple(Wc[p + 1]), (pu, Wc[p + 1])
            g_u = K.depool2_bwd(g_v, st['masksB'][p - 1], sizes[p - 1][0], sizes[p - 1][1], (ul, vl), (pu[0], pu[2]),
                                (pu[1] - pu[0], pu[3] - pu[2]))
            skip[p] = (g_u, pu)          # u_p = c_{p+1} + pool_p (p < P) or pool_P itself: gradient of pool_p over window pu
            g_c = g_u                    # ... and of c_{p+1}
            if ae_level:                 # + d mean((h - h_hat)^2) / d c_{p+1} = 2 c / n over the whole map (pool_p's share cancels)
                g_c = torch.zeros_like(st['ae_c'])
                g_c[:, pu[0]:pu[1], pu[2]:pu[3]].copy_(g_u)
                K.ae_grad_add(g_c, st['ae_c'], self.sums[8:10])
        # contracting path, bottleneck towards the input
        g_in = None
        for p in range(P, 0, -1):
            lay = self.down[p - 1]
            hp, wp = sizes[p - 1]
            pooled = st['pools'][p - 1]
            gs, pu = skip[p]
            if g_in is None:             # level P: only the up path reaches pool_P
                g_pool = torch.zeros_like(pooled)
                g_pool[:, pu[0]:pu[1], pu[2]:pu[3]].copy_(gs)
            else:
                g_pool = g_in            # dgrad of conv_{p+1} already carries the skip part (epilogue addend)
            g_a = K.pool2_relu_bwd(g_pool, pooled, st['masksA'][p - 1], hp, wp, zmask=st['zmasks'][p - 1])
            if self.keep_grads is not None:          # parity tooling: gradient w.r.t. the pre-rectifier output of conv_p
                self.keep_grads[p] = (g_pool, g_a)
            pad = geo.padding if (p == 1 and geo.padding > 0) else 1
            if p - 1 == geo.n_pool:
                xs = [(st['h'], geo.h_pad), (st['pools'][p - 2], st['pools'][p - 2].shape[3])]
            elif p == 1:
                xs = [(st['x0'], 16)]
            else:
                xs = [(st['pools'][p - 2], st['pools'][p - 2].shape[3])]
            if p == 1 and self.border_once and 0 in self._dwin:
                # the network input is zero outside the image (padding 100): only the output pixels whose 3x3 field meets it --
                # the y-dependent window of level 1 -- contribute to dW; the bias sum still runs over the whole map
                hl, wl, OH, OW = self._dwin[0]
                self._wgrad(lay, g_a[:, hl:hl + OH, wl:wl + OW].contiguous(), xs, (hl, wl), pad, bias_from=g_a)
            else:
                self._wgrad(lay, g_a, xs, (0, 0), pad)
            if p > 1:
                gs_prev, pw = skip[p - 1]
                buf = torch.zeros_like(st['pools'][p - 2])
                buf[:, pw[0]:pw[1], pw[2]:pw[3]].copy_(gs_prev)       # data movement only: the skip gradient in place
                g_in = self._dgrad(lay, g_a, (1, 1, hp, wp), addend=buf)
        s = self.sums
        self.last_loss = s     # device tensor; loss = s0/s1 + lmb*s2/s3 (+ the dice term): K.loss_from_sums
        return [lay.grad for lay in self.layers()]

    def _ae_sums(self):
        self.sums[8:10].zero_()
        K.sq_sum(self.st['ae_c'], self.sums[8:10])

    def loss_value(self):
        return K.loss_from_sums(self.sums.cpu(), self.lmb, self.terms, self.ae_h)

    def update(self):
        if self.optimizer == 'adam':
            K.adam_advance(self.adam_state, self.lr, self.beta1, self.beta2)
            for lay in self.layers():
                if not hasattr(lay, 'mom'):
                    lay.mom, lay.mom_b = torch.zeros_like(lay.w), torch.zeros_like(lay.b)
                K.adam_pack(lay.w, lay.mom, lay.acc, lay.b, lay.mom_b, lay.acc_b, lay.grad, lay.wb, lay.wt, 9, lay.cin_pad, lay.bias_col,
                            lay.ci0, lay.ci_t, self.adam_state, self.beta1, self.beta2, self.adam_eps, g_rstride=lay.g_rstride)
            return
        for lay in self.layers():
            K.rmsprop_pack(lay.w, lay.acc, lay.b, lay.acc_b, lay.grad, lay.wb, lay.wt, 9, lay.cin_pad, lay.bias_col,
                           lay.ci0, lay.ci_t, self.lr, self.rho, self.eps, g_rstride=lay.g_rstride)

    def step(self, h_bf16, y, target, noise_main=None, noise_mask=None, world=None):
        self.forward(h_bf16, y, noise_main, noise_mask)
        self.backward(target, world)
        if world is not None:           # per-rank gradients already carry the global denominators: plain sum
            world.allreduce_sum(self._grad_flat)
        self.update()
        return self.last_loss

    def step_dp(self, h_bf16, y, target, noise_main=None, noise_mask=None, world=None):
        """Data-parallel step with the gradient all-reduce bucketed (BUCKET_BYTES) and launched under backward: bucket k
        is reduced on NCCL's stream while the layers of bucket k+1.. are still being differentiated; the update waits
        for the last bucket only.  Same result as `step(..., world=world)` (same sums, same order inside NCCL)."""
        from . import _lib
        self.forward(h_bf16, y, noise_main, noise_mask)
        # the persistent conv kernels take every SM and all of its shared memory, so NCCL's CTAs mostly start in the gaps
        # between them; DP_RESERVED_SMS > 0 leaves SMs free while gradients are in flight (measured: not a win, see above)
        prev = _lib.load().iiseg_reserve_sms(self.DP_RESERVED_SMS if (world is not None and world.size > 1) else 0)
        try:
            self.backward(target, world, overlap=True)
        finally:
            _lib.load().iiseg_reserve_sms(prev)
        for w in self._works:
            w.wait()                   # stream-level wait: the update kernels queue behind the reductions
        self._works, self._dp_world = [], None
        self.update()
        return self.last_loss

    def dp_info(self):
        return {'buckets': len(self._buckets), 'bucket_bytes': [int((hi - lo) * 4) for lo, hi, _ in self._buckets],
                'gradient_bytes': int(self._grad_flat.numel() * 4), 'overlap': 'buckets all-reduced asynchronously under backward'}

    def step_graphed(self, h_bf16, y, target, noise_main=None, noise_mask=None):
        """Single-device `step` as one CUDA graph replay (~170 launches per step otherwise go through the host one by
        one).  The first call with a new shape signature runs eagerly (it also sizes the workspaces), the second
        captures and replays, later ones copy the inputs into the graph's static buffers and replay: every call is
        exactly one training step."""
        ins = [h_bf16, y, target, noise_main, noise_mask]
        key = tuple((tuple(t.shape), t.dtype) if t is not None else None for t in ins)
        # lr / sigma / lmb / rho / eps are kernel arguments passed by value, i.e. frozen into a captured graph: they are
        # part of the graph's identity, and a change (the reference anneals lr every epoch, train_dae.py:424) drops the
        # stale graph and captures a new one on the next call
        hp = (float(self.lr), float(self.sigma), float(self.lmb), float(self.rho), float(self.eps))
        ent = self._graphs.get(key)
        if ent is not None and ent != 'warm' and ent[2] != hp:
            ent = self._graphs[key] = 'warm'
        if ent is None:
            self._graphs[key] = 'warm'
            return self.step(*ins)
        if ent == 'warm':
            bufs = [t.clone() if t is not None else None for t in ins]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):             # capture executes nothing
                self.step(*bufs)
            ent = self._graphs[key] = (g, bufs, hp)
        g, bufs, _ = ent
        for b_, t in zip(bufs, ins):
            if t is not None:
                b_.copy_(t)
        g.replay()
        return self.last_loss

    def grads_lasagne(self):
        """Gradients in the reference's parameter layout (for tests): [(dW (Cout,Cin,3,3), db (Cout,)), ...]."""
        out = []
        for lay, splits in zip(self.layers(), self._splits):
            g = lay.grad[:, :3 * lay.g_rstride].reshape(lay.cout_pad, 3, lay.g_rstride)[:, :, :3 * lay.cin_pad]
            g = g.reshape(lay.cout_pad, 3, 3, lay.cin_pad)
            parts, c0 = [], 0
            for real, padded in splits:
                parts.append(g[:lay.cout, :, :, c0:c0 + real])
                c0 += padded
            out += [torch.cat(parts, dim=3).permute(0, 3, 1, 2).contiguous(), lay.grad[:lay.cout, lay.bias_col].clone()]
        return out


# ---------------------------------------------------------------------------
# The host loop of the reference's train() (train_dae.py:351-457): epochs, validation, learning-rate annealing,
# patience, best / last checkpoints in the reference's positional .npz layout.
# ---------------------------------------------------------------------------
def validate(trainer, h_bf16, y, target, noise_mask=None):
    """`val_fn` of train_dae.py:338: [test_loss, test_jacc (2, C) float32, test_mse_loss].  deterministic=True switches the
    main GaussianNoiseLayer off, but NOT the DePool2D sub-graphs (layers/mylayers.py:91-93): for a DAE with noise > 0 the
    reference's validation masks are noised as well, one draw per DePool2D (`noise_mask` [P, B, C, H, W]; tests/golden/
    ref_train_noise.npz); with noise == 0 everything is deterministic."""
    from .functions import MetricsAccumulator, jaccard_from_cm
    logits = trainer.forward(h_bf16, y, None, noise_mask)
    K.loss_grad(logits, target, trainer.C, trainer.lmb, trainer.sums, passes=1, terms=trainer.terms)          # loss sums only
    if trainer.ae_h:          # test_loss += squared_error_L(h_test, h_hat_test).mean() (train_dae.py:319)
        trainer._ae_sums()
    loss = K.loss_from_sums(trainer.sums.cpu(), trainer.lmb, trainer.terms, trainer.ae_h)
    B, _, H, W = y.shape
    p = torch.empty((B, trainer.C, H, W), dtype=torch.float32, device=y.device)
    K.softmax_nchw(logits, trainer.C, p)
    acc = MetricsAccumulator(B, trainer.C, y.device)
    K.metrics_accumulate(p, acc.cm, acc.counts, acc.sqerr, onehot=target, void_label=trainer.C)
    se = acc.sqerr.sum(0).cpu().numpy()
    return loss, jaccard_from_cm(acc.cm.sum(0).cpu().numpy()), float(se[0] / se[1])


def train(dataset, segm_net, learning_rate=0.005, lr_anneal=1.0, weight_decay=1e-4, num_epochs=500, max_patience=100,
          optimizer='rmsprop', training_loss=['squared_error'], batch_size=[10, 1, 1], ae_h=False, dae_dict_updates={},
          data_augmentation={}, savepath=None, loadpath=None, resume=False, train_from_0_255=False, lmb=1,
          full_im_ft=False, train_iter=None, val_iter=None, fcn_params=None, dae_params=None, weights_path=None,
          seed=0, verbose=True):
    """Same signature as the reference's train() (train_dae.py:54-60) plus in-memory iterators / checkpoints.  Built for
    the benchmark configuration (kind='standard', trackind unpool, rmsprop, crossentropy + squared_error).  As in the
    reference, `weight_decay` only enters the experiment name (train_dae.py:15,55,91)."""
    import os
    import time
    import numpy as np
    from .data_loader import load_data
    from .helpers import build_experiment_name
    from .iterative_inference import DAE_DICT_DEFAULTS
    from .models.fcn8 import buildFCN8
    from ._packing import load_npz_params
    dae_dict = dict(DAE_DICT_DEFAULTS)
    dae_dict['path_weights'] = ''
    dae_dict.update(dae_dict_updates)
    if optimizer not in ('rmsprop', 'adam'):
        raise ValueError('Unknown optimizer')          # train_dae.py:331
    if dae_dict['kind'] != 'standard' or dae_dict['unpool_type'] != 'trackind' or segm_net not in ('fcn8', 'densenet'):
        raise NotImplementedError('B200 train step: kind=standard, unpool_type=trackind, segmentation_net in (fcn8, densenet)')
    if not training_loss or set(training_loss) - set(K.LOSS_TERMS):
        raise NotImplementedError('B200 train step: training_loss terms among %s (squared_error_h is not built)' % sorted(K.LOSS_TERMS))
    if ae_h and 'pool' not in dae_dict['concat_h'][-1]:
        raise ValueError('Plug&Play version needs concat_h to be different than input')          # train_dae.py:179-180
    exp_name = build_experiment_name(segm_net, training_loss=training_loss, data_aug=bool(data_augmentation),
                                     learning_rate=learning_rate, lr_anneal=lr_anneal, weight_decay=weight_decay,
                                     optimizer=optimizer, ae_h=ae_h, **dae_dict)
    if savepath is None:
        raise ValueError('A saving directory must be specified')
    loadpath_init = os.path.join(loadpath, dataset, exp_name) if loadpath is not None else None      # train_dae.py:96
    exp_name += '_ft' if full_im_ft else ''
    savepath = os.path.join(savepath, dataset, exp_name)
    os.makedirs(savepath, exist_ok=True)
    if train_iter is None or val_iter is None:
        train_iter = load_data(dataset, data_augmentation, one_hot=True, batch_size=batch_size, which_set='train')
        val_iter = load_data(dataset, {}, one_hot=True, batch_size=batch_size, which_set='val')
    n_classes = train_iter.non_void_nclasses
    if segm_net == 'fcn8':           # train_dae.py:156-163
        fcn = buildFCN8(train_iter.data_shape[0], None, n_classes=n_classes, layer=dae_dict['concat_h'] + [dae_dict['layer']],
                        path_weights=os.path.join(weights_path or '', dataset, 'fcn8_model.npz'), params=fcn_params)
        padding, hkey = 100, dae_dict['concat_h'][-1]
    else:                            # train_dae.py:164-168: FC-DenseNet103 conditioning, no padding
        from .models.FCDenseNet import build_fcdensenet
        fcn = build_fcdensenet(None, dae_dict['concat_h'], train_iter.data_shape[0], n_classes,
                               weight_path=os.path.join(weights_path or '', dataset, 'DenseNet103', 'weights', 'FC-DenseNet103_weights.npz'),
                               params=fcn_params)
        padding, hkey = 0, dae_dict['concat_h'][-1] + '_bf16'
    fnet = fcn[0].net
    if dae_params is None:
        if resume:
            # resume / full_im_ft read <loadpath>/<dataset>/<exp_name>/dae_model_best.npz (train_dae.py:96,186-187)
            dae_params = load_npz_params(os.path.join(loadpath_init or savepath, 'dae_model_best.npz'))
        else:
            from . import synthetic
            dae_params = synthetic.synthetic_dae_params(n_classes, fcn[0].output_shape[1], seed=seed, n_filters=dae_dict['n_filters'],
                                                        concat_h=tuple(dae_dict['concat_h']), additional_pool=dae_dict['additional_pool'])
    tr = DAETrainer(n_classes, fcn[0].output_shape[1], padding, dae_params, concat_h=tuple(dae_dict['concat_h']),
                    n_filters=dae_dict['n_filters'], additional_pool=dae_dict['additional_pool'],
                    learning_rate=learning_rate, noise=dae_dict['noise'], lmb=lmb, optimizer=optimizer,
                    training_loss=tuple(training_loss), ae_h=ae_h)
    gen = torch.Generator(device=tr.dev).manual_seed(seed)
    say = print if verbose else (lambda *a, **k: None)

    def batch(it):
        X, L = it.next()
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(tr.dev)
        Ld = torch.from_numpy(np.ascontiguousarray(L, dtype=np.float32)).to(tr.dev)
        out = fnet.forward(Xd, want=(dae_dict['concat_h'][-1], 'probs_dimshuffle'))
        y = Ld[:, :n_classes].contiguous() if dae_dict['from_gt'] else out['probs_dimshuffle']
        return out[hkey], y, Ld

    err_train, err_valid, jacc_val_arr, mse_val_arr = [], [], [], []
    patience, best_err_val = 0, None
    for epoch in range(num_epochs):
        t0 = time.time()
        cost = 0.0
        for _ in range(train_iter.nbatches):
            h, y, Ld = batch(train_iter)
            nm = nk = None
            if tr.sigma > 0:
                nm = torch.randn(y.shape, device=tr.dev, generator=gen)
                nk = torch.randn((tr.geo.total,) + tuple(y.shape), device=tr.dev, generator=gen)      # one draw per DePool2D
            tr.step_graphed(h, y, Ld, nm, nk)
            cost += tr.loss_value()
        err_train.append(cost / train_iter.nbatches)
        cv, jv, mv = 0.0, 0, 0.0
        for _ in range(val_iter.nbatches):
            h, y, Ld = batch(val_iter)
            nk = torch.randn((tr.geo.total,) + tuple(y.shape), device=tr.dev, generator=gen) if tr.sigma > 0 else None
            c, j, m = validate(tr, h, y, Ld, nk)
            cv += c; jv = jv + j; mv += m
        err_valid.append(cv / val_iter.nbatches)
        with np.errstate(divide='ignore', invalid='ignore'):
            jacc_val_arr.append(float(np.mean(jv[0, :] / jv[1, :])))
        mse_val_arr.append(mv / val_iter.nbatches)
        out_str = 'EPOCH %i: Avg epoch training cost train %f, cost val %f, jacc val %f, mse val % f took %f s' % (
            epoch, err_train[epoch], err_valid[epoch], jacc_val_arr[epoch], mse_val_arr[epoch], time.time() - t0)
        say(out_str)
        with open(os.path.join(savepath, 'output.log'), 'a') as f:
            f.write(out_str + '\n')
        tr.lr = float(tr.lr * lr_anneal)                 # lr.set_value(lr * lr_anneal), train_dae.py:424
        arrays = [np.asarray(a.cpu()) for a in tr.params()]
        if epoch == 0:
            best_err_val = err_valid[epoch]
        elif err_valid[epoch] < best_err_val:
            best_err_val = err_valid[epoch]
            patience = 0
            np.savez(os.path.join(savepath, 'dae_model_best.npz'), *arrays)
            np.savez(os.path.join(savepath, 'dae_errors_best.npz'), err_train, err_valid, jacc_val_arr, mse_val_arr)
        else:
            patience += 1
            np.savez(os.path.join(savepath, 'dae_model_last.npz'), *arrays)
            np.savez(os.path.join(savepath, 'dae_errors_last.npz'), err_train, err_valid, jacc_val_arr, mse_val_arr)
        if patience == max_patience or epoch == num_epochs - 1:
            if loadpath is not None:
                import shutil
                dst = os.path.join(loadpath, dataset, exp_name)
                if os.path.abspath(dst) != os.path.abspath(savepath):
                    shutil.copytree(savepath, dst, dirs_exist_ok=True)
            say(' Training Done !')
            break
    return {'err_train': err_train, 'err_valid': err_valid, 'jacc_val': jacc_val_arr, 'mse_val': mse_val_arr,
            'savepath': savepath, 'trainer': tr}


def main():
    """The reference's CLI (train_dae.py:460-505); `-train_dict` / `-dae_dict` / `-data_augmentation` take dict literals."""
    import argparse
    from .iterative_inference import _flag, _literal_dict
    parser = argparse.ArgumentParser(description='DAE training')
    parser.add_argument('-dataset', type=str, default='camvid', help='Dataset.')
    parser.add_argument('-segmentation_net', type=str, default='fcn8', help='Segmentation network.')
    parser.add_argument('-train_dict', type=_literal_dict,
                        default={'learning_rate': 0.001, 'lr_anneal': 0.99, 'weight_decay': 0.0001, 'num_epochs': 500,
                                 'max_patience': 100, 'optimizer': 'rmsprop', 'batch_size': [10, 10, 10],
                                 'training_loss': ['crossentropy', 'squared_error'], 'lmb': 1, 'full_im_ft': False},
                        help='Training configuration')
    parser.add_argument('-dae_dict', type=_literal_dict,
                        default={'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0.5,
                                 'concat_h': ['pool4'], 'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1,
                                 'additional_pool': 2, 'temperature': 1.0, 'path_weights': '', 'layer': 'probs_dimshuffle',
                                 'exp_name': 'flip_final_', 'bn': 0}, help='DAE kind and parameters')
    parser.add_argument('-data_augmentation', type=_literal_dict,
                        default={'crop_size': (224, 224), 'horizontal_flip': 0.5, 'fill_mode': 'constant'},
                        help='Dictionary of data augmentation to be used')
    parser.add_argument('-train_from_0_255', type=_flag, default=False)
    parser.add_argument('-savepath', type=str, default='./iiseg_out/')
    parser.add_argument('-loadpath', type=str, default='./iiseg_models/')
    parser.add_argument('-weights_path', type=str, default='./iiseg_models/')
    args = parser.parse_args()
    train(dataset=args.dataset, segm_net=args.segmentation_net, dae_dict_updates=args.dae_dict,
          data_augmentation=args.data_augmentation, train_from_0_255=args.train_from_0_255, resume=False,
          savepath=args.savepath, loadpath=args.loadpath, weights_path=args.weights_path, **args.train_dict)


if __name__ == '__main__':
    main()

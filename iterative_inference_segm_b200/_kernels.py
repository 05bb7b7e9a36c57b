"""Torch-tensor front end of the C ABI: shape bookkeeping + pointer passing.

PyTorch is only the allocator and the stream here; every function launches the
library's sm_100a kernels on torch's current CUDA stream (so calls can be
captured by torch.cuda.graph).  Nothing in this module computes on the CPU.
"""
import ctypes as C

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise _lib.IisegError('%s must be a contiguous CUDA %s tensor' % (name, dtype))


def require_device():
    if not torch.cuda.is_available():
        raise _lib.IisegError('no CUDA device: libiiseg has no CPU fallback')
    _lib.call('iiseg_device_check', torch.cuda.current_device())


def pad_channels(c, conv_input=True, narrow=False):
    """Channel padding rule of the kernels: conv inputs are multiples of 64 (one 128-byte TMA row);
    small outputs (score maps / logits) are 16.  `narrow`: a <=16-channel input of a 3x3 conv may be
    padded to 16 instead (32-byte rows, one K=16 MMA per tap; the halo-tile kernel, bf16 variant)."""
    if (not conv_input or narrow) and c <= 16:
        return 16
    return (c + 63) // 64 * 64


# ---- layout ---------------------------------------------------------------
def pack_nchw(src, cpad, out=None, split=False):
    """NCHW fp32 -> NHWC bf16 [N,H,W,cpad]; split: [N,H,W,2*cpad] = the (hi | lo) bf16 pair."""
    _chk(src, F32, 'src')
    N, Cc, H, W = src.shape
    cs = 2 * cpad if split else cpad
    if out is None:
        out = torch.empty((N, H, W, cs), dtype=BF16, device=src.device)
    assert tuple(out.shape) == (N, H, W, cs)
    _lib.call('iiseg_pack_nchw_f32_to_nhwc_bf16', _ptr(src), _ptr(out), N, Cc, H, W, cpad, int(split), _stream())
    return out


def unpack_nhwc(src, c, out=None, split=False):
    N, H, W, cpad = src.shape
    if split:
        cpad //= 2
    if out is None:
        out = torch.empty((N, c, H, W), dtype=F32, device=src.device)
    if src.dtype == BF16:
        _lib.call('iiseg_unpack_nhwc_bf16_to_nchw_f32', _ptr(src), _ptr(out), N, c, H, W, cpad, int(split), _stream())
    else:
        _chk(src, F32, 'src')
        _lib.call('iiseg_unpack_nhwc_f32_to_nchw_f32', _ptr(src), _ptr(out), N, c, H, W, cpad, _stream())
    return out


def widen_nhwc(src, out, split=False):
    """NHWC bf16 [N,H,W,C] (split: the (hi | lo) pair [N,H,W,2C]) -> NHWC fp32 `out` [N,H,W,C] (= hi + lo)."""
    _chk(src, BF16, 'src')
    _chk(out, F32, 'out')
    Cc = out.shape[3]
    assert tuple(src.shape[:3]) == tuple(out.shape[:3]) and src.shape[3] == (2 * Cc if split else Cc) and Cc % 8 == 0
    _lib.call('iiseg_widen_nhwc_bf16_to_f32', _ptr(src), _ptr(out), out.shape[0] * out.shape[1] * out.shape[2], Cc, int(bool(split)), _stream())
    return out


# ---- convolution ------------------------------------------------------------
def conv_out_size(H, W, R, S, pad):
    return H + 2 * pad - R + 1, W + 2 * pad - S + 1


def conv2d(src0, weight, bias, R, S, pad, relu, src1=None, addend=None, window=None, out=None,
           out_f32=False, addend_off=(0, 0), pooled=None, pool_mask=None, split=False, update=None,
           out_slice=None, pool_zmask=None, depool=None, depool_out=None, addend_pair_hi=False, out_strided=None,
           src_pair_hi=False, post_affine=None, weight_npack=None):
    """src0/src1: NHWC bf16; weight: bf16 [Cout, R*S*(C0+C1)]; bias fp32 [Cout].
    window = (oh0, ow0, OH, OW) selects the output window (default: all).

    split=True is the fp32-accurate variant: every activation tensor (src0, src1, addend, out,
    pooled) carries the (hi | lo) bf16 pair of an fp32 map, 2*C channels per pixel, and `weight` is
    packed by _packing.pack_conv(split=True) as [Cout, R*S*3*(C0+C1)] = per tap (W_hi | W_hi | W_lo):
    the loader walks (hi, lo, hi) views of the sources, so the unchanged bf16 tensor-core loop
    accumulates hi*hi + lo*hi + hi*lo in fp32.

    update = dict(y, y_bf16, active, norm_acc, step, C[, y_split]): the 16-channel logits conv with the softmax
    tail and the iterative-inference update fused in its epilogue (iiseg_conv_desc.upd_*); nothing is
    returned, y / y_bf16 / norm_acc are updated in place.  y_split: y_bf16 carries the (hi | lo) pair of y.

    out_strided = (dest, stride, (h0, w0)): output pixel (oh, ow) goes to dest[n, oh*stride + h0, ow*stride + w0]
    (dest: [N,DH,DW,cm*Cout], bf16 or fp32 per out_f32) and the addend, if any, is read at the same stride from
    `addend_off` -- one output phase of a transposed convolution (iiseg_conv_desc.out_stride).  Returns dest.

    weight_npack: the same 16-row 3x3 filter re-packed by _packing.pack_npack16; the library then runs the N-packed kernel
    where it applies (iiseg_conv_desc.weight_npack).

    post_affine = (scale, shift): fp32 [Cout] vectors applied after the rectifier, before pool / store (a folded
    deterministic BatchNormLayer, iiseg_conv_desc.post_scale); (scale, shift, mean): ((x - mean) * scale) + shift with one
    rounding per operation -- lasagne's (x - mean) * (gamma * inv_std) + beta (iiseg_conv_desc.post_mean).

    src_pair_hi: src0 is a (hi | lo) pair tensor [N,H,W,2*C0] of which a plain bf16 conv reads the hi halves.

    addend_pair_hi: the (bf16, non-split) conv adds the hi halves of a (hi | lo) pair tensor `addend`
    [N,AH,AW,2*Cout] (iiseg_conv_desc.addend_cs): the skip-sum of the mixed-precision expanding path.

    depool = (mask, H, W, u_origin): `src0` is the POOLED tensor u (a dense window whose element (0,0) sits at pooled
    position u_origin) and the conv runs on the virtual map DePool2D(u, mask) of size HxW, expanded inside the
    kernel's loader (iiseg_conv_desc.depool_mask) -- same result as unpool2(...) followed by conv2d(...).

    depool_out = (v, mask, (h0, w0), (ph0, pw0)): the conv's output u is not stored; its epilogue writes
    v = DePool2D(u, mask) restricted to the window tensor `v` [N,VH,VW,Cout] whose element (0,0) is full-resolution
    pixel (h0, w0); `mask` [N,H2,W2,Cout/8] is the tie mask over the pooled grid, on which this launch's output pixel
    (0,0) sits at (ph0, pw0).  Returns v.

    out_slice = (stack, c_off): fp32 output written as channels [c_off, c_off + Cout) of the wider fp32
    NHWC tensor `stack` [N,OH,OW,Cs] (the DenseNet stack: ConcatLayer without a copy)."""
    _chk(src0, BF16, 'src0')
    _chk(weight, BF16, 'weight')
    _chk(bias, F32, 'bias')
    N, H, W, C0 = src0.shape
    if depool is not None:
        dp_mask, H, W, dp_origin = depool
        _chk(dp_mask, torch.int32, 'depool.mask')
        assert src1 is None and not split and tuple(dp_mask.shape) == (N, H // 2, W // 2, C0 // 8), (tuple(dp_mask.shape), H, W, C0)
    C1 = 0
    if src1 is not None:
        _chk(src1, BF16, 'src1')
        assert src1.shape[:3] == src0.shape[:3]
        C1 = src1.shape[3]
    Cout = weight.shape[0]
    cm = 2 if (split and not out_f32) else 1     # channel multiplier of out / pooled / addend tensors
    if src_pair_hi:
        assert not split and src1 is None
        C0 //= 2
    if split:
        C0 //= 2
        C1 //= 2
        assert weight.shape[1] == R * S * 3 * (C0 + C1), (tuple(weight.shape), R, S, C0, C1)
    else:
        assert weight.shape[1] == R * S * (C0 + C1), (tuple(weight.shape), R, S, C0, C1)
    fOH, fOW = conv_out_size(H, W, R, S, pad)
    oh0, ow0, OH, OW = window if window is not None else (0, 0, fOH, fOW)
    if pooled is not None:     # fused 2x2 max-pool (+ mask): the full-resolution output is never written
        _chk(pooled, BF16, 'pooled')
        dense = tuple(pooled.shape) == (N, OH // 2, OW // 2, cm * Cout)
        if not dense:   # a window of a larger pooled tensor: even origin, pooled rows [oh0/2, oh0/2 + OH/2)
            assert pooled.shape[0] == N and pooled.shape[3] == cm * Cout and oh0 % 2 == 0 and ow0 % 2 == 0
            assert oh0 // 2 + OH // 2 <= pooled.shape[1] and ow0 // 2 + OW // 2 <= pooled.shape[2]
        pool_hw = (0, 0) if dense else (pooled.shape[1], pooled.shape[2])
        if pool_mask is not None:
            _chk(pool_mask, torch.int32, 'pool_mask')
            assert tuple(pool_mask.shape) == tuple(pooled.shape[:3]) + (Cout // 8,)
        assert out is None
    elif update is not None:
        _chk(update['y'], F32, 'update.y')
        _chk(update['y_bf16'], BF16, 'update.y_bf16')
        _chk(update['norm_acc'], torch.int64, 'update.norm_acc')
        assert Cout == 16 and out is None and addend is None and not split
        assert tuple(update['y'].shape) == (N, update['C'], OH, OW) and tuple(update['y_bf16'].shape[:3]) == (N, OH, OW)
        assert update['norm_acc'].numel() == N
    elif depool_out is not None:
        dpo_v, dpo_mask, dpo_org, dpo_porg = depool_out
        _chk(dpo_v, BF16, 'depool_out.v')
        _chk(dpo_mask, torch.int32, 'depool_out.mask')
        assert out is None and not split and not out_f32 and Cout % 64 == 0
        assert dpo_v.shape[0] == N and dpo_v.shape[3] == Cout and dpo_mask.shape[0] == N and dpo_mask.shape[3] == Cout // 8
        assert dpo_porg[0] + OH <= dpo_mask.shape[1] and dpo_porg[1] + OW <= dpo_mask.shape[2]
    elif out_strided is not None:
        dest, ostride, (oh_0, ow_0) = out_strided
        _chk(dest, F32 if out_f32 else BF16, 'out_strided.dest')
        assert out is None and dest.shape[0] == N and dest.shape[3] == cm * Cout and ostride >= 2
        assert oh_0 + (OH - 1) * ostride < dest.shape[1] and ow_0 + (OW - 1) * ostride < dest.shape[2]
    elif out_slice is not None:
        stack, c_off = out_slice
        _chk(stack, F32, 'out_slice.stack')
        assert out_f32 and out is None and tuple(stack.shape[:3]) == (N, OH, OW) and c_off + Cout <= stack.shape[3] and c_off % 4 == 0
    elif out is None:
        out = torch.empty((N, OH, OW, cm * Cout), dtype=F32 if out_f32 else BF16, device=src0.device)
    else:
        _chk(out, F32 if out_f32 else BF16, 'out')
        assert tuple(out.shape) == (N, OH, OW, cm * Cout), (tuple(out.shape), (N, OH, OW, cm * Cout))
    addend_f32 = addend is not None and addend.dtype == F32     # hoisted fp32 term (see iiseg_conv_desc.addend_f32)
    addend_cs = 0
    if addend is not None:
        _chk(addend, F32 if addend_f32 else BF16, 'addend')
        if addend_pair_hi:
            assert not split and not addend_f32 and addend.shape[3] == 2 * Cout
            addend_cs = 2 * Cout
        assert addend.shape[0] == N and addend.shape[3] == (Cout if addend_f32 else (2 * Cout if addend_pair_hi else cm * Cout))
        ast = out_strided[1] if out_strided is not None else 1
        assert addend_off[0] + (OH - 1) * ast < addend.shape[1] and addend_off[1] + (OW - 1) * ast < addend.shape[2]
    d = _lib.ConvDesc(N=N, H=H, W=W, weight=weight.data_ptr(), bias=bias.data_ptr(),
                      Cout=Cout, R=R, S=S, pad=pad, oh0=oh0, ow0=ow0, OH=OH, OW=OW,
                      out=out.data_ptr() if out is not None else None,
                      addend=addend.data_ptr() if addend is not None else None,
                      pooled=pooled.data_ptr() if pooled is not None else None,
                      pool_mask=pool_mask.data_ptr() if pool_mask is not None else None,
                      pool_zmask=pool_zmask.data_ptr() if pool_zmask is not None else None,
                      pool_H=pool_hw[0] if pooled is not None else 0, pool_W=pool_hw[1] if pooled is not None else 0,
                      AH=addend.shape[1] if addend is not None else 0, AW=addend.shape[2] if addend is not None else 0,
                      ah0=addend_off[0], aw0=addend_off[1], addend_f32=int(addend_f32), addend_cs=addend_cs,
                      relu=int(bool(relu)), split=int(bool(split and not out_f32)), out_f32=int(bool(out_f32)))
    if weight_npack is not None:
        _chk(weight_npack, BF16, 'weight_npack')
        assert tuple(weight_npack.shape) == (39 * 16, 64) and Cout == 16 and R == 3 and S == 3
        d.weight_npack = weight_npack.data_ptr()
    if post_affine is not None:
        _chk(post_affine[0], F32, 'post_affine.scale')
        _chk(post_affine[1], F32, 'post_affine.shift')
        assert post_affine[0].numel() == Cout == post_affine[1].numel()
        d.post_scale, d.post_shift = post_affine[0].data_ptr(), post_affine[1].data_ptr()
        if len(post_affine) == 3:          # (gamma * inv_std, beta, mean): ((x - mean) * scale) + shift, rounded like lasagne's expression
            _chk(post_affine[2], F32, 'post_affine.mean')
            assert post_affine[2].numel() == Cout
            d.post_mean = post_affine[2].data_ptr()
    if out_strided is not None:
        d.out, d.out_stride, d.out_H, d.out_W, d.out_h0, d.out_w0 = dest.data_ptr(), ostride, dest.shape[1], dest.shape[2], oh_0, ow_0
    if depool_out is not None:
        d.depool_out, d.depool_out_mask = dpo_v.data_ptr(), dpo_mask.data_ptr()
        d.depool_out_VH, d.depool_out_VW, d.depool_out_h0, d.depool_out_w0 = dpo_v.shape[1], dpo_v.shape[2], dpo_org[0], dpo_org[1]
        d.depool_out_H2, d.depool_out_W2, d.depool_out_ph0, d.depool_out_pw0 = dpo_mask.shape[1], dpo_mask.shape[2], dpo_porg[0], dpo_porg[1]
    if depool is not None:
        d.depool_mask = dp_mask.data_ptr()
        d.depool_UH, d.depool_UW, d.depool_h0, d.depool_w0 = src0.shape[1], src0.shape[2], dp_origin[0], dp_origin[1]
    if out_slice is not None:
        d.out, d.out_cs = out_slice[0].data_ptr() + 4 * out_slice[1], out_slice[0].shape[3]
    if update is not None:
        d.upd_y, d.upd_y_bf16 = update['y'].data_ptr(), update['y_bf16'].data_ptr()
        d.upd_active = update['active'].data_ptr() if update.get('active') is not None else None
        d.upd_norm_acc = update['norm_acc'].data_ptr()
        ysp = bool(update.get('y_split'))
        stp = update['step']          # a python float, or a 1-element fp32 CUDA tensor read by the kernel at run time
        if isinstance(stp, torch.Tensor):
            _chk(stp, F32, 'update.step')
            d.upd_step, d.upd_step_dev = 0.0, stp.data_ptr()
        else:
            d.upd_step = float(stp)
        d.upd_C, d.upd_split = int(update['C']), int(ysp)
        d.upd_cpad = int(update['y_bf16'].shape[3]) // (2 if ysp else 1)
    # the channel-concatenated source views, in K order: (pointer, channels, channels per pixel in memory)
    srcs = [(src0, C0)] + ([(src1, C1)] if src1 is not None else [])
    views = []
    if split:
        for half in (0, 1, 0):        # hi, lo, hi  against  W_hi, W_hi, W_lo
            views += [(t.data_ptr() + 2 * half * c, c, 2 * c) for t, c in srcs]
    else:
        views = [(t.data_ptr(), c, 2 * c if src_pair_hi else 0) for t, c in srcs]
    assert len(views) <= _lib.MAX_SRC
    for i, (ptr, c, cs) in enumerate(views):
        d.src[i], d.C[i], d.Cs[i] = ptr, c, cs
    _lib.call('iiseg_conv2d_fwd', C.byref(d), _stream())
    if depool_out is not None:
        return dpo_v
    if out_strided is not None:
        return dest
    return out if pooled is None else pooled


# ---- pool / unpool ----------------------------------------------------------
def maxpool2(x, with_mask, pooled=None, mask=None):
    _chk(x, BF16, 'x')
    N, H, W, Cc = x.shape
    if pooled is None:
        pooled = torch.empty((N, H // 2, W // 2, Cc), dtype=BF16, device=x.device)
    if with_mask and mask is None:
        mask = torch.empty((N, H // 2, W // 2, Cc // 8), dtype=torch.int32, device=x.device)
    _lib.call('iiseg_maxpool2_mask_fwd', _ptr(x), _ptr(pooled), _ptr(mask) if with_mask else C.c_void_p(0),
              N, H, W, Cc, _stream())
    return (pooled, mask) if with_mask else pooled


def unpool2(u, mask, H, W, out=None, u_origin=(0, 0), window=None, split=False):
    """DePool2D into the HxW pre-pool map.  `u` is a dense window of the pooled map starting at pooled
    position `u_origin`; `window` = (h0, w0, OH, OW) restricts the output (default: the whole map).
    split=True: u and out carry (hi | lo) pairs; split=2: u is a pair tensor, out its gated hi halves (plain bf16)."""
    _chk(u, BF16, 'u')
    _chk(mask, torch.int32, 'mask')
    N, UH, UW, Cu = u.shape
    Cc = Cu // 2 if split else Cu          # split: u / out carry (hi | lo) pairs, the mask has Cc channels
    assert tuple(mask.shape) == (N, H // 2, W // 2, Cc // 8), (tuple(mask.shape), H, W, Cc)
    h0, w0, OH, OW = window if window is not None else (0, 0, H, W)
    Co = Cc if split == 2 else Cu
    if out is None:
        out = torch.empty((N, OH, OW, Co), dtype=BF16, device=u.device)
    assert tuple(out.shape) == (N, OH, OW, Co)
    _lib.call('iiseg_unpool2_mask_window_fwd', _ptr(u), _ptr(mask), _ptr(out), N, H, W, Cc, UH, UW,
              u_origin[0], u_origin[1], OH, OW, h0, w0, int(split), _stream())
    return out


# ---- transposed conv on 16-channel fp32 maps --------------------------------
def deconv16(x, weight, bias, k, stride, window=None, addend=None, addend_off=(0, 0), out=None):
    _chk(x, F32, 'x')
    _chk(weight, F32, 'weight')
    _chk(bias, F32, 'bias')
    N, H, W, Cc = x.shape
    assert Cc == 16 and tuple(weight.shape) == (k, k, 16, 16)
    fH, fW = (H - 1) * stride + k, (W - 1) * stride + k
    oh0, ow0, OH, OW = window if window is not None else (0, 0, fH, fW)
    if out is None:
        out = torch.empty((N, OH, OW, 16), dtype=F32, device=x.device)
    AH = AW = 0
    if addend is not None:
        _chk(addend, F32, 'addend')
        AH, AW = addend.shape[1], addend.shape[2]
    d = _lib.DeconvDesc(x=x.data_ptr(), N=N, H=H, W=W, weight=weight.data_ptr(), bias=bias.data_ptr(),
                        k=k, stride=stride, oh0=oh0, ow0=ow0, OH=OH, OW=OW,
                        addend=addend.data_ptr() if addend is not None else None, AH=AH, AW=AW,
                        ah0=addend_off[0], aw0=addend_off[1], out=out.data_ptr())
    _lib.call('iiseg_deconv2d_fwd', C.byref(d), _stream())
    return out


# ---- context-module DAE: small-channel (dilated) 3x3 conv on planar fp32 ---------
def ctx_conv(x, weight, bias, dil, out, relu=True, origin=(0, 0), check=False, out_origin=(0, 0), size=None, addend=None,
             active=None, tail=None, in_nhwc=False, out_nhwc=False):
    """models/contextmod_dae.py:72-103 on the CUDA cores (csrc/contextmod.cu).  x planar fp32 [N,Cin,Hin,Win], or with
    `in_nhwc` channels-last [N,Hin,Win,pad4(Cin)]; weight / bias are HOST numpy float32 arrays ([Cin,3,3,Cout] and [Cout]):
    they ride in the kernel's parameter block.  `size` = (OH, OW) computed; output pixel (oh, ow), tap (r, s) reads
    x[.., oh + origin[0] + r*dil, ow + origin[1] + s*dil] (`check`: zero outside x).  out planar fp32 [N,Cout,Hout,Wout]
    (`out_nhwc`: [N,Hout,Wout,pad4(Cout)], as is `addend` then), written at `out_origin`; or with tail=(w2 [Cout,C2], b2 [C2])
    the 1x1 linear conv is applied to the rectified result and out is the fp32 NHWC16 logits tensor [N,OH,OW,16]."""
    import numpy as np
    _chk(x, F32, 'x')
    _chk(out, F32, 'out')
    Cin = weight.shape[0]
    pad4 = lambda c: (c + 3) & ~3          # noqa: E731
    if in_nhwc:
        N, Hin, Win, cpi = x.shape
        assert cpi == pad4(Cin), (tuple(x.shape), Cin)
    else:
        N, c_, Hin, Win = x.shape
        assert c_ == Cin, (tuple(x.shape), Cin)
    assert isinstance(weight, np.ndarray) and weight.dtype == np.float32 and weight.flags['C_CONTIGUOUS'] \
        and weight.shape[1:3] == (3, 3), 'weight: host float32 [Cin,3,3,Cout]'
    Cout = weight.shape[3]
    assert isinstance(bias, np.ndarray) and bias.dtype == np.float32 and bias.shape == (Cout,)
    if tail is None:
        if out_nhwc:
            assert out.shape[0] == N and out.shape[3] == pad4(Cout), (tuple(out.shape), Cout)
            Hout, Wout = out.shape[1], out.shape[2]
        else:
            assert out.shape[0] == N and out.shape[1] == Cout
            Hout, Wout = out.shape[2], out.shape[3]
        OH, OW = size if size is not None else (Hout, Wout)
        w2 = b2 = None
        C2 = 0
    else:
        w2, b2 = tail
        assert w2.dtype == np.float32 and b2.dtype == np.float32 and w2.flags['C_CONTIGUOUS'] and w2.shape[0] == Cout and b2.shape == (w2.shape[1],)
        C2 = w2.shape[1]
        assert out.shape[0] == N and out.shape[3] == 16
        OH, OW = out.shape[1], out.shape[2]
        Hout, Wout = OH, OW
    if addend is not None:
        _chk(addend, F32, 'addend')
        assert tuple(addend.shape) == ((N, OH, OW, pad4(Cout)) if out_nhwc else (N, Cout, OH, OW))
    d = _lib.CtxConvDesc(in_=x.data_ptr(), N=N, Cin=Cin, Hin=Hin, Win=Win, in_h0=origin[0], in_w0=origin[1], check=int(check),
                         dil=dil, out=out.data_ptr(), Cout=Cout, Hout=Hout, Wout=Wout, out_h0=out_origin[0],
                         out_w0=out_origin[1], OH=OH, OW=OW, weight=weight.ctypes.data, bias=bias.ctypes.data,
                         addend=addend.data_ptr() if addend is not None else None,
                         active=active.data_ptr() if active is not None else None, relu=int(relu),
                         weight2=w2.ctypes.data if w2 is not None else None, bias2=b2.ctypes.data if b2 is not None else None,
                         C2=C2, in_nhwc=int(in_nhwc), out_nhwc=int(out_nhwc), stream=torch.cuda.current_stream().cuda_stream)
    _lib.call('iiseg_ctx_conv', C.byref(d))
    return out


# ---- softmax / update ---------------------------------------------------------
def _scalar(v):
    """(by-value float, device pointer) of a kernel scalar: python floats go by value; a 1-element fp32 CUDA tensor is
    read by the kernel when it runs (so a captured graph is not tied to the value)."""
    if isinstance(v, torch.Tensor):
        _chk(v, F32, 'scalar')
        return C.c_float(0.0), C.c_void_p(v.data_ptr())
    return C.c_float(float(v)), C.c_void_p(0)


def update_blocks(H, W):
    return _lib.load().iiseg_update_blocks(H, W)


def softmax_nchw(logits, C_, p_out, y_bf16=None, split=False):
    _chk(logits, F32, 'logits')
    N, H, W, c16 = logits.shape
    assert c16 == 16
    _chk(p_out, F32, 'p_out')
    cpad = y_bf16.shape[3] // (2 if split else 1) if y_bf16 is not None else 0
    _lib.call('iiseg_softmax_nchw', _ptr(logits), _ptr(p_out), _ptr(y_bf16), N, C_, H, W, cpad, int(split), _stream())
    return p_out


def softmax_update(logits, y, y_bf16, active, norm_partial, step, p_out=None, split=False):
    _chk(logits, F32, 'logits')
    _chk(y, F32, 'y')
    N, C_, H, W = y.shape
    assert tuple(logits.shape) == (N, H, W, 16)
    cpad = y_bf16.shape[3] // (2 if split else 1) if y_bf16 is not None else 0
    sv, sd = _scalar(step)
    _lib.call('iiseg_softmax_update', _ptr(logits), _ptr(y), _ptr(y_bf16), _ptr(p_out), _ptr(active),
              _ptr(norm_partial), N, C_, H, W, cpad, sv, sd, int(split), _stream())


def softmax_grad(logits, y, grad):
    _chk(logits, F32, 'logits')
    _chk(y, F32, 'y')
    _chk(grad, F32, 'grad')
    N, C_, H, W = y.shape
    _lib.call('iiseg_softmax_grad', _ptr(logits), _ptr(y), _ptr(grad), N, C_, H, W, _stream())
    return grad


def norm_finalize_fixed(norm_acc, norm, active, n_exec, H, W, eps):
    """norm / active / n_exec step from the fixed-point accumulator of the fused conv epilogue."""
    _chk(norm_acc, torch.int64, 'norm_acc')
    N = norm.shape[0]
    ev, ed = _scalar(eps)
    _lib.call('iiseg_norm_finalize_fixed', _ptr(norm_acc), _ptr(norm), _ptr(active), _ptr(n_exec), N, H, W,
              ev, ed, _stream())


def norm_finalize(norm_partial, norm, active, n_exec, H, W, eps):
    N = norm.shape[0]
    ev, ed = _scalar(eps)
    _lib.call('iiseg_norm_finalize', _ptr(norm_partial), _ptr(norm), _ptr(active), _ptr(n_exec), N, H, W,
              ev, ed, _stream())


# ---- FC-DenseNet103 streaming kernels -------------------------------------------
def bn_relu_pack(stack, C, out, c0=0, stats=None, gamma=None, beta=None, relu=True, split=False):
    """Channels [c0, c0+C) of the fp32 NHWC `stack` -> zero-padded bf16 NHWC `out`, through BatchNorm with the batch
    statistics `stats` = (mean, inv_std) + gamma/beta and rectify; stats=None: plain convert.
    split: out is [N,H,W,2*Cpad], the (hi | lo) bf16 pair of each fp32 value."""
    _chk(stack, F32, 'stack')
    _chk(out, BF16, 'out')
    N, H, W, Cs = stack.shape
    cpad = out.shape[3] // (2 if split else 1)
    assert tuple(out.shape[:3]) == (N, H, W) and cpad >= C
    mean, inv_std = stats if stats is not None else (None, None)
    _lib.call('iiseg_bn_relu_pack', _ptr(stack), N, H, W, Cs, c0, C, _ptr(mean), _ptr(inv_std), _ptr(gamma), _ptr(beta),
              int(bool(relu)), _ptr(out), cpad, int(bool(split)), _stream())
    return out


def channel_stats(stack, c0, nC, mean, inv_std, scratch, eps=1e-4):
    """Batch statistics of channels [c0, c0+nC) of `stack` into mean[c0:c0+nC], inv_std[c0:c0+nC] (fp32 vectors)."""
    _chk(stack, F32, 'stack')
    N, H, W, Cs = stack.shape
    need = _lib.load().iiseg_channel_stats_chunks(N, H, W) * nC * 2
    assert scratch.dtype == torch.float64 and scratch.numel() >= need
    _lib.call('iiseg_channel_stats', _ptr(stack), N, H, W, Cs, c0, nC, C.c_float(eps), _ptr(scratch), C_void(mean, c0),
              C_void(inv_std, c0), _stream())


def C_void(vec, off):
    _chk(vec, F32, 'vector')
    return C.c_void_p(vec.data_ptr() + 4 * off)


def maxpool2_f32(x, C_, out):
    _chk(x, F32, 'x')
    _chk(out, F32, 'out')
    N, H, W, Cs = x.shape
    assert tuple(out.shape[:3]) == (N, H // 2, W // 2)
    _lib.call('iiseg_maxpool2_f32', _ptr(x), N, H, W, Cs, C_, _ptr(out), out.shape[3], _stream())
    return out


def deconv_interleave(phases, C_, crop, out):
    """phases[py][px]: dense fp32 [N,H+1,W+1,Cp]; out: fp32 [N,OH,OW,Cs], first C_ channels written."""
    p00 = phases[0][0]
    N, H1, W1, Cp = p00.shape
    for row in phases:
        for t in row:
            _chk(t, F32, 'phase')
            assert tuple(t.shape) == (N, H1, W1, Cp)
    _chk(out, F32, 'out')
    _lib.call('iiseg_deconv_interleave', _ptr(phases[0][0]), _ptr(phases[0][1]), _ptr(phases[1][0]), _ptr(phases[1][1]),
              N, H1 - 1, W1 - 1, Cp, C_, crop[0], crop[1], _ptr(out), out.shape[1], out.shape[2], out.shape[3], _stream())
    return out


# ---- DAE training step kernels ------------------------------------------------------
def noise_pack(y, noise, sigma, cpad, out=None, split=False):
    """bf16 NHWC [N,H,W,cpad] of y + sigma*noise (NCHW fp32 inputs; noise=None: plain pack); split: the (hi | lo) pair."""
    _chk(y, F32, 'y')
    N, Cc, H, W = y.shape
    if noise is not None:
        _chk(noise, F32, 'noise')
        assert noise.shape == y.shape
    if out is None:
        out = torch.empty((N, H, W, (2 if split else 1) * cpad), dtype=BF16, device=y.device)
    _lib.call('iiseg_noise_pack', _ptr(y), _ptr(noise), C.c_float(sigma), _ptr(out), N, Cc, H, W, cpad, int(bool(split)), _stream())
    return out


LOSS_TERMS = {'crossentropy': 1, 'squared_error': 2, 'dice': 4}          # `terms` bits of iiseg_loss_grad_terms


def loss_grad(logits, target, n_classes, lmb, sums, dlogits=None, passes=3, terms=3):
    """Loss sums and d loss / d logits of the terms train_dae.py:278-294 adds up (`terms`: OR of LOSS_TERMS values)."""
    _chk(logits, F32, 'logits')
    _chk(target, F32, 'target')
    N, H, W, c16 = logits.shape
    assert c16 == 16 and tuple(target.shape) == (N, n_classes + 1, H, W) and sums.dtype == torch.float64
    assert sums.numel() >= (8 if terms & 4 else 4)
    if dlogits is None:
        dlogits = torch.empty((N, H, W, 16), dtype=BF16, device=logits.device)
    _lib.call('iiseg_loss_grad_terms', _ptr(logits), _ptr(target), N, n_classes, H, W, C.c_float(lmb), int(terms), _ptr(sums), _ptr(dlogits),
              passes, _stream())
    return dlogits


def sq_sum(x, sums2):
    """sums2[0] += sum x^2, sums2[1] += x.numel() (the ae_h term of train_dae.py:317-319; x bf16, sums2 fp64 [2])."""
    _chk(x, BF16, 'x')
    assert sums2.dtype == torch.float64 and sums2.numel() == 2 and sums2.is_contiguous()
    _lib.call('iiseg_sq_sum', _ptr(x), x.numel(), _ptr(sums2), _stream())


def broadcast_image(t):
    """t[1:] = t[0] for a contiguous batched CUDA tensor (bytes per image a multiple of 16)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise _lib.IisegError('broadcast_image: contiguous CUDA tensor expected')
    _lib.call('iiseg_broadcast_image', _ptr(t), t[0].numel() * t.element_size(), t.shape[0], _stream())


def add_bf16(a, b):
    """a + b (bf16 tensors of one shape; fp32 sum, one rounding): a skip sum outside a conv epilogue."""
    _chk(a, BF16, 'a')
    _chk(b, BF16, 'b')
    assert a.shape == b.shape
    out = torch.empty_like(a)
    _lib.call('iiseg_add_bf16', _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
    return out


def ae_grad_add(g, c, sums2):
    """g += 2 c / sums2[1]: the gradient of mean(c^2) over the (global) element count held on the device."""
    _chk(g, BF16, 'g')
    _chk(c, BF16, 'c')
    assert g.shape == c.shape and sums2.dtype == torch.float64 and sums2.numel() == 2
    _lib.call('iiseg_ae_grad_add', _ptr(g), _ptr(c), g.numel(), _ptr(sums2), _stream())


def loss_from_sums(s, lmb, terms=3, ae_h=False):
    """The scalar loss from the (host) sums of `loss_grad` (+ `sq_sum` in s[8:10] with ae_h)."""
    loss = float(s[8] / s[9]) if ae_h else 0.0
    if terms & 1:
        loss += float(s[0] / s[1])
    if terms & 4:
        loss += float(-(2.0 * s[4] + 1.0) / (s[5] + s[6] + 1.0))
    if terms & 2:
        loss += lmb * float(s[2] / s[3])
    return loss


def depool2_bwd(gv, mask, H, W, v_origin, u_origin, u_size):
    """gv: [N,VH,VW,C] window of the unpooled-map gradient at full-resolution origin v_origin; returns the
    pooled-map gradient over the window (u_origin, u_size)."""
    _chk(gv, BF16, 'gv')
    _chk(mask, torch.int32, 'mask')
    N, VH, VW, Cc = gv.shape
    assert tuple(mask.shape) == (N, H // 2, W // 2, Cc // 8)
    gu = torch.empty((N, u_size[0], u_size[1], Cc), dtype=BF16, device=gv.device)
    _lib.call('iiseg_depool2_bwd', _ptr(gv), _ptr(mask), _ptr(gu), N, H, W, Cc, VH, VW, v_origin[0], v_origin[1],
              u_size[0], u_size[1], u_origin[0], u_origin[1], _stream())
    return gu


def pool2_relu_bwd(gpool, pooled, mask, H, W, zmask=None):
    _chk(gpool, BF16, 'gpool')
    _chk(pooled, BF16, 'pooled')
    _chk(mask, torch.int32, 'mask')
    N, H2, W2, Cc = pooled.shape
    assert (H2, W2) == (H // 2, W // 2) and gpool.shape == pooled.shape and tuple(mask.shape) == (N, H2, W2, Cc // 8)
    ga = torch.empty((N, H, W, Cc), dtype=BF16, device=gpool.device)
    _lib.call('iiseg_pool2_relu_bwd', _ptr(gpool), _ptr(pooled), _ptr(mask), _ptr(zmask), _ptr(ga), N, H, W, Cc, _stream())
    return ga


def transpose_shift(x, C_, origin, size, shift, out, row0, c0=0, nshift=1, shift_rows=0):
    """out[row0 + c, p] = x[n, origin+o+shift, c0 + c] (zero outside the map), p = flat (n, oh, ow) over `size`;
    nshift > 1 also writes the copies s at rows + s*shift_rows holding the same matrix at flat pixel p + s."""
    _chk(x, BF16, 'x')
    _chk(out, BF16, 'out')
    N, H, W, Cs = x.shape
    assert out.dim() == 2 and row0 + (nshift - 1) * shift_rows + C_ <= out.shape[0] and out.shape[1] >= N * size[0] * size[1]
    _lib.call('iiseg_transpose_shift', _ptr(x), N, H, W, Cs, c0, C_, origin[0], origin[1], size[0], size[1], shift[0], shift[1],
              _ptr(out), C.c_longlong(out.shape[1]), C.c_longlong(row0), nshift, C.c_longlong(shift_rows), _stream())


def wgrad_gemm(gT, xT, cin_pad, groups, slabs, out_ld, out=None):
    """Weight gradient of a conv without a 9-fold im2col:
        G[co][t*cin_pad + ci] = sum_k gT[co][k] * xT[row_t + ci][k + koff_t]      for (row_t, koff_t) = groups[t]
    (columns beyond xT's width read as zero; koff_t % 8 == 0) -> fp32 [Cg, out_ld], first len(groups)*cin_pad columns written.
    gT [Cg, K], xT [rows, K]: bf16, K = slabs * slab, slab % 64 == 0 -- the pixel axis of a zero-padded grid.
    One launch of the tcgen05 conv kernel: a 1x1 conv whose batch images are the K slabs (split-K) and whose
    output-channel groups are the taps (`w_groups`); `iiseg_sum_slabs` adds the slab partials in order."""
    _chk(gT, BF16, 'gT')
    _chk(xT, BF16, 'xT')
    M, Kt = gT.shape
    taps = len(groups)
    assert xT.shape[1] == Kt and Kt % (64 * slabs) == 0 and out_ld >= taps * cin_pad and out_ld % 4 == 0
    slab = Kt // slabs
    if out is not None:           # a persistent destination (a view of the trainer's flat gradient buffer)
        _chk(out, F32, 'out')
        assert tuple(out.shape) == (M, out_ld)
    part = out.view(1, M, out_ld) if (out is not None and slabs == 1) else torch.empty((slabs, M, out_ld), dtype=F32, device=gT.device)
    zero = torch.zeros((taps * cin_pad,), dtype=F32, device=gT.device)
    d = _lib.ConvDesc(N=slabs, H=1, W=M, weight=xT.data_ptr(), bias=zero.data_ptr(), Cout=taps * cin_pad, R=1, S=1, pad=0,
                      oh0=0, ow0=0, OH=1, OW=M, out=part.data_ptr(), relu=0, out_f32=1, out_cs=out_ld,
                      src_image_stride=slab, weight_ld=Kt, w_koff=slab, w_groups=taps, w_rows_total=xT.shape[0])
    d.src[0], d.C[0], d.Cs[0] = gT.data_ptr(), slab, Kt
    for t, (row, k) in enumerate(groups):
        d.w_group_row[t], d.w_group_koff[t] = row, k
    _lib.call('iiseg_conv2d_fwd', C.byref(d), _stream())
    if slabs == 1:
        return part.view(M, out_ld)
    if out is None:
        out = torch.empty((M, out_ld), dtype=F32, device=gT.device)
    _lib.call('iiseg_sum_slabs', _ptr(part), _ptr(out), slabs, C.c_longlong(M * out_ld), _stream())
    return out


def bias_grad(g, out, col, chunks=592):
    """out[c][col] = sum over the pixels of the bf16 NHWC gradient g[..., c]  (out: fp32 [Cg, ld])."""
    _chk(g, BF16, 'g')
    _chk(out, F32, 'out')
    Cg = g.shape[-1]
    P = g.numel() // Cg
    chunks = max(1, min(chunks, P // 64))
    scratch = torch.empty((chunks, Cg), dtype=F32, device=g.device)
    _lib.call('iiseg_bias_grad', _ptr(g), C.c_longlong(P), Cg, _ptr(scratch), chunks, out.data_ptr() + 4 * col, out.shape[1], _stream())


def last_conv_plan():
    """(kernel, BN, KB) of this thread's last conv launch: kernel 0 per-tap, 1 CTA pair, 2 halo tile."""
    k, bn, kb = C.c_int(-1), C.c_int(0), C.c_int(0)
    _lib.load().iiseg_last_conv_plan(C.byref(k), C.byref(bn), C.byref(kb))
    return k.value, bn.value, kb.value


def gemm_nt_splitk(A, Bm, slabs):
    """G[m][n] = sum_k A[m][k] * Bm[n][k] for row-major bf16 A [M, K], Bm [Nn, K] (K = slabs * slab, slab % 64 == 0),
    fp32 result [M, Nn].  Runs on the tcgen05 conv kernel as a 1x1 conv whose batch images are the K slabs (split-K: each
    slab yields a partial product, `iiseg_sum_slabs` adds them in order)."""
    _chk(A, BF16, 'A')
    _chk(Bm, BF16, 'Bm')
    M, Kt = A.shape
    Nn = Bm.shape[0]
    assert Bm.shape[1] == Kt and Kt % (64 * slabs) == 0 and (Nn == 16 or Nn % 64 == 0)
    slab = Kt // slabs
    part = torch.empty((slabs, 1, M, Nn), dtype=F32, device=A.device)
    zero = torch.zeros((Nn,), dtype=F32, device=A.device)
    d = _lib.ConvDesc(N=slabs, H=1, W=M, weight=Bm.data_ptr(), bias=zero.data_ptr(), Cout=Nn, R=1, S=1, pad=0, oh0=0, ow0=0,
                      OH=1, OW=M, out=part.data_ptr(), relu=0, out_f32=1,
                      src_image_stride=slab if slabs > 1 else 0, weight_ld=Kt if slabs > 1 else 0, w_koff=slab if slabs > 1 else 0)
    d.src[0], d.C[0], d.Cs[0] = A.data_ptr(), slab, Kt
    _lib.call('iiseg_conv2d_fwd', C.byref(d), _stream())
    if slabs == 1:
        return part.view(M, Nn)
    out = torch.empty((M, Nn), dtype=F32, device=A.device)
    _lib.call('iiseg_sum_slabs', _ptr(part), _ptr(out), slabs, C.c_longlong(M * Nn), _stream())
    return out


def rmsprop_pack(w, acc, b, acc_b, g, wb, wt, taps, cin_pad, bias_col, ci0, ci_t, lr, rho, eps, g_rstride=0):
    _chk(w, F32, 'w'); _chk(acc, F32, 'acc'); _chk(b, F32, 'b'); _chk(acc_b, F32, 'acc_b'); _chk(g, F32, 'g'); _chk(wb, BF16, 'wb')
    Cout = w.shape[0]
    assert w.numel() == Cout * taps * cin_pad == acc.numel() == wb.numel() and g.shape[0] == Cout
    co_pad = 0
    if wt is not None:
        _chk(wt, BF16, 'wt')
        assert wt.shape[0] == ci_t and wt.shape[1] % taps == 0
        co_pad = wt.shape[1] // taps
    _lib.call('iiseg_rmsprop_pack', _ptr(w), _ptr(acc), _ptr(b), _ptr(acc_b), _ptr(g), _ptr(wb), _ptr(wt), Cout, taps, cin_pad,
              g.shape[1], g_rstride, bias_col, ci0, ci_t, co_pad, C.c_float(lr), C.c_float(rho), C.c_float(eps), _stream())


def adam_pack(w, m, v, b, m_b, v_b, g, wb, wt, taps, cin_pad, bias_col, ci0, ci_t, state, beta1, beta2, eps, g_rstride=0):
    """lasagne.updates.adam on the layouts of `rmsprop_pack`; `state` = fp32 [2] device tensor (t, a_t) kept by `adam_advance`."""
    for t_, n_ in ((w, 'w'), (m, 'm'), (v, 'v'), (b, 'b'), (m_b, 'm_b'), (v_b, 'v_b'), (g, 'g'), (state, 'state')):
        _chk(t_, F32, n_)
    _chk(wb, BF16, 'wb')
    Cout = w.shape[0]
    assert w.numel() == Cout * taps * cin_pad == m.numel() == v.numel() == wb.numel() and g.shape[0] == Cout and state.numel() == 2
    co_pad = 0
    if wt is not None:
        _chk(wt, BF16, 'wt')
        assert wt.shape[0] == ci_t and wt.shape[1] % taps == 0
        co_pad = wt.shape[1] // taps
    _lib.call('iiseg_adam_pack', _ptr(w), _ptr(m), _ptr(v), _ptr(b), _ptr(m_b), _ptr(v_b), _ptr(g), _ptr(wb), _ptr(wt), Cout, taps, cin_pad,
              g.shape[1], g_rstride, bias_col, ci0, ci_t, co_pad, state.data_ptr() + 4, C.c_float(beta1), C.c_float(beta2), C.c_float(eps), _stream())


def adam_advance(state, lr, beta1, beta2):
    """t <- t + 1, a_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), on the device (once per training step)."""
    _chk(state, F32, 'state')
    _lib.call('iiseg_adam_advance', _ptr(state), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), _stream())


# ---- metrics ------------------------------------------------------------------
def metrics_accumulate(y, cm, counts, sqerr, onehot=None, labels=None, active=None, void_label=-1):
    _chk(y, F32, 'y')
    N, C_, H, W = y.shape
    if onehot is not None:
        _chk(onehot, F32, 'onehot')
        assert tuple(onehot.shape) == (N, C_ + 1, H, W), tuple(onehot.shape)
    if labels is not None:
        _chk(labels, torch.int32, 'labels')
    _chk(cm, torch.int64, 'cm')
    _chk(counts, torch.int64, 'counts')
    _chk(sqerr, torch.float64, 'sqerr')
    _lib.call('iiseg_metrics_accumulate', _ptr(y), _ptr(onehot), _ptr(labels), _ptr(active), _ptr(cm),
              _ptr(counts), _ptr(sqerr), N, C_, H, W, void_label, _stream())


def onehot_to_labels(onehot, labels=None):
    _chk(onehot, F32, 'onehot')
    N, C1, H, W = onehot.shape
    if labels is None:
        labels = torch.empty((N, H, W), dtype=torch.int32, device=onehot.device)
    _lib.call('iiseg_onehot_to_labels', _ptr(onehot), _ptr(labels), N, C1, H, W, _stream())
    return labels

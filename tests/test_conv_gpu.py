"""GPU parity of the tcgen05 implicit-GEMM conv and the SIMT transposed conv against a plain
PyTorch fp32 reference on the SAME bf16-rounded operands (so the only differences are the fp32
accumulation order and the final bf16 rounding of the output: tolerance 2^-7 relative, stated
below).  Shapes cover ragged tiles, the pad=100 first layer, the dual-source (concat) loader,
the fused skip-sum, the cropped fp32 output window, 1x1 and 7x7 filters."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32
CASES = [
    (1, 8, 16, 64, 0, 64, 1, 0, 0, 0, None, 0),
    (1, 8, 16, 64, 0, 256, 1, 0, 1, 0, None, 0),
    (2, 37, 45, 64, 0, 128, 3, 1, 1, 0, None, 0),
    (2, 17, 21, 256, 0, 512, 3, 1, 1, 0, None, 0),
    (2, 17, 21, 128, 64, 256, 3, 1, 1, 0, None, 0),
    (2, 33, 29, 128, 0, 64, 3, 1, 0, 1, None, 0),
    (1, 24, 32, 64, 0, 64, 3, 100, 1, 0, None, 0),            # halo-tile kernel, split precision, 64-channel K blocks
    (2, 30, 40, 16, 0, 64, 3, 100, 1, 0, None, 0),            # ... 16-channel K blocks (y / RGB inputs)
    (2, 41, 37, 16, 0, 64, 3, 1, 0, 0, (3, 5, 30, 28), 0),    # ... linear, windowed
    (1, 40, 56, 64, 0, 16, 3, 1, 0, 0, (5, 7, 24, 32), 1),
    (1, 17, 21, 128, 0, 256, 7, 0, 1, 0, None, 0),
    (1, 11, 15, 256, 0, 16, 1, 0, 1, 0, None, 1),
    (4, 139, 169, 64, 0, 128, 3, 1, 1, 0, None, 0),
    # 16-channel source: 32-byte rows, SWIZZLE_32B, one K=16 MMA per tap (halo-tile kernel)
    (2, 37, 45, 16, 0, 64, 3, 1, 1, 0, None, 0),
    (1, 24, 32, 16, 0, 64, 3, 100, 1, 0, None, 0),
    (2, 30, 41, 16, 0, 128, 3, 1, 0, 0, (3, 5, 20, 30), 0),
]


@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_conv_matches_fp32_reference(cuda, case):
    from iterative_inference_segm_b200 import _kernels as K
    N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32 = case
    torch.manual_seed(0)
    Cin = C0 + C1
    x = torch.randn(N, Cin, H, W, device=cuda).to(torch.bfloat16)
    Wt = (torch.randn(Cout, Cin, R, R, device=cuda) / (Cin * R * R) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=cuda)
    xn = x.permute(0, 2, 3, 1).contiguous()
    src0 = xn[..., :C0].contiguous()
    src1 = xn[..., C0:].contiguous() if C1 else None
    Wk = Wt.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    fOH, fOW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    oh0, ow0, OH, OW = window if window else (0, 0, fOH, fOW)
    add = torch.randn(N, OH, OW, Cout, device=cuda).to(torch.bfloat16) if addend else None
    out = K.conv2d(src0, Wk, b, R, R, pad, relu, src1=src1, addend=add, window=window, out_f32=bool(out_f32))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), Wt.float(), b, padding=pad)[:, :, oh0:oh0 + OH, ow0:ow0 + OW]
    if add is not None:
        ref = ref + add.float().permute(0, 3, 1, 2)
    if relu:
        ref = torch.relu(ref)
    got = out.float().permute(0, 3, 1, 2)
    rel = 1e-4 if out_f32 else 2.0 ** -7      # fp32 out: accumulation order only; bf16 out: + output rounding
    assert bool(((got - ref).abs() <= rel * ref.abs().clamp(min=1.0)).all()), float((got - ref).abs().max())


@pytest.mark.parametrize('N,H,W,C,Cout,pad', [(2, 37, 45, 64, 128, 1), (1, 17, 21, 128, 256, 1), (1, 20, 24, 64, 64, 100),
                                               (3, 8, 10, 64, 512, 1), (2, 37, 45, 16, 64, 1), (1, 20, 24, 16, 64, 100),
                                               (2, 61, 47, 256, 64, 1)])
def test_conv_fused_pool_equals_conv_then_pool(cuda, N, H, W, C, Cout, pad):
    """The pool fused in the conv epilogue is bit-identical to conv (bf16 out) followed by the
    stand-alone pool + tie-mask kernel, including the dropped odd row/col.  Shapes are chosen so that
    the pooled and the plain launch run the same main loop (halo-tile or per-tap kernel: their fp32
    accumulation orders differ): shuffle pool on 64- and 16-channel halo tiles, smem-staged pool of
    the per-tap kernel for BN = 64 / 256."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(4)
    x = torch.randn(N, H, W, C, device=cuda).to(torch.bfloat16)
    Wk = (torch.randn(Cout, 9 * C, device=cuda) / (9 * C) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=cuda)
    full = K.conv2d(x, Wk, b, 3, 3, pad, relu=True)
    ref_p, ref_m = K.maxpool2(full, with_mask=True)
    OH, OW = full.shape[1], full.shape[2]
    pooled = torch.zeros((N, OH // 2, OW // 2, Cout), dtype=torch.bfloat16, device=cuda)
    mask = torch.zeros((N, OH // 2, OW // 2, Cout // 8), dtype=torch.int32, device=cuda)
    K.conv2d(x, Wk, b, 3, 3, pad, relu=True, pooled=pooled, pool_mask=mask)
    assert torch.equal(pooled, ref_p)
    assert torch.equal(mask, ref_m)


def test_conv_addend_window_offset(cuda):
    """Skip-sum partner read at an offset inside a larger tensor (cone-restricted expanding path)."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(3)
    N, C, Cout, H, W = 2, 64, 64, 20, 26
    x = torch.randn(N, H, W, C, device=cuda).to(torch.bfloat16)
    Wt = (torch.randn(Cout, C, 3, 3, device=cuda) / 24).to(torch.bfloat16)
    b = torch.randn(Cout, device=cuda)
    big = torch.randn(N, 40, 50, Cout, device=cuda).to(torch.bfloat16)
    win = (1, 1, 18, 24)
    out = K.conv2d(x, Wt.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), b, 3, 3, 1, relu=False, window=win,
                   addend=big, addend_off=(7, 9))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), Wt.float(), b, padding=1)[:, :, 1:19, 1:25]
    ref = ref + big[:, 7:25, 9:33].float().permute(0, 3, 1, 2)
    got = out.float().permute(0, 3, 1, 2)
    assert bool(((got - ref).abs() <= 2.0 ** -7 * ref.abs().clamp(min=1.0)).all())


def test_conv_constant_region_is_bit_constant(cuda):
    """Spatially constant input -> every interior output pixel must be bit-identical (one uniform K
    loop, zero padding by TMA fill), which the tie-inclusive pool mask relies on."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(1)
    C, Cout, H, W = 64, 64, 40, 50
    x = torch.randn(1, 1, 1, C, device=cuda).expand(1, H, W, C).to(torch.bfloat16).contiguous()
    Wk = (torch.randn(Cout, 9 * C, device=cuda) / 24).to(torch.bfloat16)
    out = K.conv2d(x, Wk, torch.zeros(Cout, device=cuda), 3, 3, 1, relu=False)
    inner = out[0, 1:-1, 1:-1].reshape(-1, Cout)
    assert bool((inner == inner[0]).all())


@pytest.mark.parametrize('k,stride,H,W,window,addend', [
    (4, 2, 11, 15, None, False), (4, 2, 11, 15, (0, 0, 24, 32), True), (16, 8, 7, 9, (24, 28, 20, 30), False)])
def test_deconv_matches_conv_transpose(cuda, k, stride, H, W, window, addend):
    """Deconv2DLayer(flip_filters=False) == conv_transpose2d with flipped kernels; fp32, tolerance 1e-4."""
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_deconv16
    torch.manual_seed(2)
    N, C = 2, 11
    x = torch.randn(N, C, H, W, device=cuda)
    Wd = torch.randn(C, C, k, k, device=cuda) / (C * (k / stride) ** 2) ** 0.5
    b = torch.randn(C, device=cuda)
    Wt, bk = pack_deconv16(Wd, b, cuda)
    xn = torch.zeros(N, H, W, 16, device=cuda)
    xn[..., :C] = x.permute(0, 2, 3, 1)
    fH, fW = (H - 1) * stride + k, (W - 1) * stride + k
    oh0, ow0, OH, OW = window if window else (0, 0, fH, fW)
    add, off = None, (0, 0)
    if addend:
        add = torch.randn(N, OH + 10, OW + 10, 16, device=cuda)
        add[..., C:] = 0
        off = (5, 5)
    out = K.deconv16(xn, Wt, bk, k, stride, window=window, addend=add, addend_off=off)
    ref = F.conv_transpose2d(x, Wd.flip(2, 3), b, stride=stride)[:, :, oh0:oh0 + OH, ow0:ow0 + OW]
    if add is not None:
        ref = ref + add[:, 5:5 + OH, 5:5 + OW, :C].permute(0, 3, 1, 2)
    got = out[..., :C].permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) < 1e-4
    assert float(out[..., C:].abs().max()) == 0.0


# ---------------------------------------------------------------------------
# fp32-accurate variant ("fp32x3"): activations and weights are (hi | lo) bf16 pairs, the conv
# accumulates hi*hi + lo*hi + hi*lo on the same tensor-core loop.  Compared against an fp64
# convolution of the ORIGINAL fp32 operands; per-product error ~2^-16 plus the tensor core's fp32
# accumulation over up to 392 MMAs (K = 6272): stated tolerance 1.5e-4 relative to max(|ref|, 1)
# (measured worst case 7.9e-5 on the 7x7 layer, ~1e-5 on the 3x3 ones; the bf16 variant's is 2^-7).
# ---------------------------------------------------------------------------
def _recon(t):
    """(hi | lo) NHWC bf16 pair tensor -> fp32 NCHW."""
    c = t.shape[3] // 2
    return (t[..., :c].float() + t[..., c:].float()).permute(0, 3, 1, 2)


SPLIT_CASES = [
    # N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32
    (2, 37, 45, 64, 0, 128, 3, 1, 1, 0, None, 0),
    (2, 17, 21, 128, 64, 256, 3, 1, 1, 0, None, 0),
    (2, 33, 29, 128, 0, 64, 3, 1, 0, 1, None, 0),
    (1, 24, 32, 64, 0, 64, 3, 100, 1, 0, None, 0),            # halo-tile kernel, split precision, 64-channel K blocks
    (2, 30, 40, 16, 0, 64, 3, 100, 1, 0, None, 0),            # ... 16-channel K blocks (y / RGB inputs)
    (2, 41, 37, 16, 0, 64, 3, 1, 0, 0, (3, 5, 30, 28), 0),    # ... linear, windowed
    (1, 40, 56, 64, 0, 16, 3, 1, 0, 0, (5, 7, 24, 32), 1),
    (1, 17, 21, 128, 0, 256, 7, 0, 1, 0, None, 0),
    (1, 11, 15, 256, 0, 16, 1, 0, 1, 0, None, 1),
]


@pytest.mark.parametrize('case', SPLIT_CASES, ids=[str(c) for c in SPLIT_CASES])
def test_conv_split_matches_fp64_reference(cuda, case):
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32 = case
    torch.manual_seed(0)
    Cin = C0 + C1
    x = torch.randn(N, Cin, H, W, device=cuda)
    Wt = torch.randn(Cout, Cin, R, R, device=cuda) / (Cin * R * R) ** 0.5
    b = torch.randn(Cout, device=cuda)
    src0 = K.pack_nchw(x[:, :C0].contiguous(), C0, split=True)
    src1 = K.pack_nchw(x[:, C0:].contiguous(), C1, split=True) if C1 else None
    Wk, bk = pack_conv(Wt, b, [(C0, C0)] + ([(C1, C1)] if C1 else []), Cout, cuda, split=True)
    fOH, fOW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    oh0, ow0, OH, OW = window if window else (0, 0, fOH, fOW)
    add_f = torch.randn(N, Cout, OH, OW, device=cuda) if addend else None
    add = K.pack_nchw(add_f, Cout, split=True) if addend else None
    out = K.conv2d(src0, Wk, bk, R, R, pad, relu, src1=src1, addend=add, window=window, out_f32=bool(out_f32),
                   split=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), Wt.double(), b.double(), padding=pad)[:, :, oh0:oh0 + OH, ow0:ow0 + OW]
    if add_f is not None:
        ref = ref + add_f.double()
    if relu:
        ref = torch.relu(ref)
    got = (out.permute(0, 3, 1, 2) if out_f32 else _recon(out)).double()
    err = ((got - ref).abs() / ref.abs().clamp(min=1.0)).max()
    assert float(err) < 1.5e-4, float(err)


@pytest.mark.parametrize('N,H,W,C,Cout,pad', [(2, 37, 45, 64, 128, 1), (1, 20, 24, 64, 64, 100), (3, 8, 10, 64, 512, 1),
                                              (2, 30, 40, 16, 64, 100), (3, 33, 47, 16, 64, 1)])
def test_conv_split_fused_pool_mask_is_exact(cuda, N, H, W, C, Cout, pad):
    """fp32x3 fused pool: the pooled pair is exactly the 2x2 max of the reconstructed fp32 (hi+lo) conv output, the tie mask
    is that of the kernel's fp32 values, and the windowed unpool gates both halves with it."""
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    from oracle import lasagne_semantics as L
    from tests.test_streaming_kernels_gpu import _mask_to_dense
    torch.manual_seed(4)
    x = torch.randn(N, C, H, W, device=cuda)
    Wt = torch.randn(Cout, C, 3, 3, device=cuda) / (9 * C) ** 0.5
    b = torch.randn(Cout, device=cuda)
    xs = K.pack_nchw(x, C, split=True)
    Wk, bk = pack_conv(Wt, b, [(C, C)], Cout, cuda, split=True)
    full = _recon(K.conv2d(xs, Wk, bk, 3, 3, pad, relu=True, split=True)).cpu()
    OH, OW = full.shape[2], full.shape[3]
    pooled = torch.zeros((N, OH // 2, OW // 2, 2 * Cout), dtype=torch.bfloat16, device=cuda)
    mask = torch.zeros((N, OH // 2, OW // 2, Cout // 8), dtype=torch.int32, device=cuda)
    K.conv2d(xs, Wk, bk, 3, 3, pad, relu=True, pooled=pooled, pool_mask=mask, split=True)
    assert torch.equal(_recon(pooled).cpu(), L.maxpool2(full))
    # The kernel decides ties on its fp32 values BEFORE they are split into pairs (kTieFp32): two values that collapse into
    # the same (hi, lo) pair are a tie of the reconstructed map but not of the fp32 one.  So the mask is a subset of the
    # reconstructed map's tie mask, never empty in a window, and equal to it wherever the reconstructed maximum is unique.
    ref_mask = L.tie_mask(full)[:, :, :2 * (OH // 2), :2 * (OW // 2)].numpy()
    dense = _mask_to_dense(mask, Cout)
    assert dense.shape == ref_mask.shape and bool(np.all(dense <= ref_mask))
    per_window = dense.reshape(N, Cout, OH // 2, 2, OW // 2, 2).sum(axis=(3, 5))
    assert bool(np.all(per_window >= 1))
    ref_window = ref_mask.reshape(N, Cout, OH // 2, 2, OW // 2, 2).sum(axis=(3, 5))
    # all-equal windows (the zero border / fully rectified windows) stay four-way ties: those are ties in fp32 as well
    zero_windows = L.maxpool2(full).numpy() == 0
    assert bool(np.all(per_window[zero_windows] == ref_window[zero_windows]))
    print('split fused pool: %d of %d windows had a tie among the (hi, lo) pairs that the fp32 values resolve'
          % (int((per_window < ref_window).sum()), per_window.size))
    u = torch.randn(N, Cout, OH // 2, OW // 2, device=cuda)
    us = K.pack_nchw(u, Cout, split=True)
    out = K.unpool2(us, mask, OH, OW, split=True)
    ur = _recon(us).cpu()
    want = torch.zeros((N, Cout, OH, OW))
    want[:, :, :2 * (OH // 2), :2 * (OW // 2)] = ur.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) * torch.from_numpy(dense)
    assert torch.equal(_recon(out).cpu(), want)


@pytest.mark.parametrize('split', [False, True])
def test_conv_hoisted_concat_half(cuda, split):
    """conv(concat(h, x)) == [conv_h(h) + b] (fp32 output, computed once) + conv_x(x) with the bracket as an
    fp32 epilogue addend: the hoist of the iteration-invariant half of DAE conv5_1
    (models/model_helpers.py:93-94).  Both forms against an fp64 convolution of the same operands."""
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    torch.manual_seed(6)
    N, H, W, Ch, Cx, Cout = 2, 18, 22, 128, 64, 256
    q = (lambda t: t) if split else (lambda t: t.to(torch.bfloat16).float())     # bf16 variant: compare on rounded operands
    h, x = q(torch.randn(N, Ch, H, W, device=cuda)), q(torch.randn(N, Cx, H, W, device=cuda))
    Wt = q(torch.randn(Cout, Ch + Cx, 3, 3, device=cuda) / (9 * (Ch + Cx)) ** 0.5)
    b = torch.randn(Cout, device=cuda)
    hs, xs = K.pack_nchw(h, Ch, split=split), K.pack_nchw(x, Cx, split=split)
    Wh, bh = pack_conv(Wt[:, :Ch], b, [(Ch, Ch)], Cout, cuda, split=split)
    Wx, _ = pack_conv(Wt[:, Ch:], b, [(Cx, Cx)], Cout, cuda, split=split)
    Wc, bc = pack_conv(Wt, b, [(Ch, Ch), (Cx, Cx)], Cout, cuda, split=split)
    hproj = K.conv2d(hs, Wh, bh, 3, 3, 1, relu=False, out_f32=True, split=split)
    assert hproj.dtype == torch.float32 and tuple(hproj.shape) == (N, H, W, Cout)
    win = (2, 4, 12, 16)
    hoisted = K.conv2d(xs, Wx, torch.zeros_like(b), 3, 3, 1, relu=True, window=win, addend=hproj, addend_off=(2, 4),
                       split=split)
    concat = K.conv2d(hs, Wc, bc, 3, 3, 1, relu=True, src1=xs, window=win, split=split)
    ref = torch.relu(F.conv2d(torch.cat([h, x], 1).double(), Wt.double(), b.double(), padding=1))[:, :, 2:14, 4:20]
    tol = 1.5e-4 if split else 2.0 ** -7
    for got in (hoisted, concat):
        g = (_recon(got) if split else got.float().permute(0, 3, 1, 2)).double()
        assert float(((g - ref).abs() / ref.abs().clamp(min=1.0)).max()) < tol


# N, H, W (pre-pool map), C, Cout, u window (ph0, pw0, UH, UW), conv window (oh0, ow0, OH, OW), addend, relu
DEPOOL_CASES = [
    (2, 40, 56, 64, 16, (0, 0, 20, 28), (0, 0, 40, 56), 0, 0),          # whole map
    (2, 41, 57, 64, 16, (0, 0, 20, 28), (0, 0, 41, 57), 0, 0),          # odd map: the last row / column has no pool window
    (2, 90, 122, 64, 64, (3, 5, 40, 52), (9, 13, 70, 96), 1, 0),        # cone window, skip-sum
    (1, 64, 80, 64, 64, (2, 2, 28, 36), (5, 7, 50, 60), 0, 1),          # odd window origin
    (3, 120, 160, 64, 16, (10, 12, 45, 60), (21, 25, 86, 116), 0, 0),   # many tiles per CTA ring
]


@pytest.mark.parametrize('case', DEPOOL_CASES, ids=[str(c) for c in DEPOOL_CASES])
def test_conv_with_fused_depool_equals_unpool_then_conv(cuda, case):
    """The DePool2D loader (iiseg_conv_desc.depool_mask) against the materialised path: same operands, same
    accumulation order -> bit-identical outputs."""
    from iterative_inference_segm_b200 import _kernels as K
    N, H, W, C, Cout, (ph0, pw0, UH, UW), (oh0, ow0, OH, OW), addend, relu = case
    torch.manual_seed(1)
    x = torch.randn(N, H, W, C, device=cuda).to(torch.bfloat16)
    x[:, 0:H - 1:2, :, :8] = x[:, 1:H:2, :, :8]                          # ties inside pool windows -> multi-bit masks
    _, mask = K.maxpool2(x, True)
    u_full = torch.randn(N, H // 2, W // 2, C, device=cuda).to(torch.bfloat16)
    u = u_full[:, ph0:ph0 + UH, pw0:pw0 + UW].contiguous()
    Wt = (torch.randn(Cout, 9 * C, device=cuda) / (9 * C) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=cuda)
    add = torch.randn(N, OH, OW, Cout, device=cuda).to(torch.bfloat16) if addend else None
    # materialised reference: unpool the window the conv needs (one pixel of halo), then the plain conv
    vh0, vw0 = max(oh0 - 1, 0), max(ow0 - 1, 0)
    vh1, vw1 = min(oh0 + OH + 1, H), min(ow0 + OW + 1, W)
    full = K.unpool2(u, mask, H, W, u_origin=(ph0, pw0), window=(2 * ph0, 2 * pw0, min(2 * UH, H - 2 * ph0), min(2 * UW, W - 2 * pw0)))
    v = torch.zeros(N, H, W, C, dtype=torch.bfloat16, device=cuda)
    v[:, 2 * ph0:2 * ph0 + full.shape[1], 2 * pw0:2 * pw0 + full.shape[2]] = full
    ref = K.conv2d(v, Wt, b, 3, 3, 1, relu=bool(relu), window=(oh0, ow0, OH, OW), addend=add, out_f32=(Cout == 16))
    got = K.conv2d(u, Wt, b, 3, 3, 1, relu=bool(relu), window=(oh0, ow0, OH, OW), addend=add, out_f32=(Cout == 16),
                   depool=(mask, H, W, (ph0, pw0)))
    torch.cuda.synchronize()
    assert vh0 >= 2 * ph0 and vw0 >= 2 * pw0 and vh1 <= min(2 * (ph0 + UH), H) + (H % 2) and vw1 <= min(2 * (pw0 + UW), W) + (W % 2)
    assert torch.equal(got, ref), float((got.float() - ref.float()).abs().max())


# N, conv input map H x W, Cin, Cout, conv window (oh0, ow0, OH, OW), v window in the 2x map (vh0, vw0, VH, VW), odd (2H+1 x 2W+1 map)
DEPOOL_OUT_CASES = [
    (2, 17, 21, 128, 256, (0, 0, 17, 21), (0, 0, 34, 42), 0),            # CTA-pair kernel, whole maps
    (2, 17, 21, 128, 256, (0, 0, 17, 21), (0, 0, 35, 43), 1),            # odd unpooled map: trailing row / column stay zero
    (2, 27, 35, 256, 128, (1, 1, 25, 33), (2, 3, 50, 64), 0),            # cone windows (up_conv3-like), pair<128>
    (3, 46, 60, 128, 64, (3, 2, 40, 55), (7, 5, 78, 108), 0),            # halo-tile kernel (up_conv2-like), skip-sum
    (1, 30, 40, 64, 64, (2, 4, 26, 30), (4, 8, 52, 59), 1),
]


@pytest.mark.parametrize('case', DEPOOL_OUT_CASES, ids=[str(c) for c in DEPOOL_OUT_CASES])
def test_conv_with_depool_epilogue_equals_conv_then_unpool(cuda, case):
    """iiseg_conv_desc.depool_out: DePool2D written by the producing conv's epilogue == conv followed by the unpool kernel,
    bit for bit (same accumulation, same mask selection)."""
    from iterative_inference_segm_b200 import _kernels as K
    N, H, W, Cin, Cout, (oh0, ow0, OH, OW), (vh0, vw0, VH, VW), odd = case
    torch.manual_seed(2)
    FH, FW = 2 * H + odd, 2 * W + odd                                   # the unpooled (pre-pool) map
    xfull = torch.randn(N, FH, FW, Cout, device=cuda).to(torch.bfloat16)
    xfull[:, 0:2 * H:2, :, :16] = xfull[:, 1:2 * H:2, :, :16]           # ties -> multi-bit masks
    _, mask = K.maxpool2(xfull, True)                                    # [N, H, W, Cout/8]
    x = torch.randn(N, H, W, Cin, device=cuda).to(torch.bfloat16)
    Wt = (torch.randn(Cout, 9 * Cin, device=cuda) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=cuda)
    add = torch.randn(N, H, W, Cout, device=cuda).to(torch.bfloat16)
    u = K.conv2d(x, Wt, b, 3, 3, 1, relu=False, window=(oh0, ow0, OH, OW), addend=add, addend_off=(oh0, ow0))
    ref = K.unpool2(u, mask, FH, FW, u_origin=(oh0, ow0), window=(vh0, vw0, VH, VW))
    v = torch.zeros(N, VH, VW, Cout, dtype=torch.bfloat16, device=cuda)
    got = K.conv2d(x, Wt, b, 3, 3, 1, relu=False, window=(oh0, ow0, OH, OW), addend=add, addend_off=(oh0, ow0),
                   depool_out=(v, mask, (vh0, vw0), (oh0, ow0)))
    torch.cuda.synchronize()
    assert torch.equal(got, ref), float((got.float() - ref.float()).abs().max())


def test_split_halo_kernel_is_selected(cuda):
    """The split-precision 3x3 convs with Cout <= 128 run on the halo-tile kernel (kernel id 2): DAE conv1_1 / FCN8 conv1_1
    (16-channel K blocks) and FCN8 conv1_2 (64 -> 64)."""
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    for C in (16, 64):
        x = K.pack_nchw(torch.randn(1, C, 40, 48, device=cuda), C, split=True)
        Wk, bk = pack_conv(torch.randn(64, C, 3, 3, device=cuda), torch.zeros(64, device=cuda), [(C, C)], 64, cuda, split=True)
        pooled = torch.empty((1, 20, 24, 128), dtype=torch.bfloat16, device=cuda)
        K.conv2d(x, Wk, bk, 3, 3, 1, relu=True, pooled=pooled, split=True)
        assert K.last_conv_plan() == (2, 64, C), K.last_conv_plan()


def test_mixed_precision_plumbing(cuda):
    """The three kernel-level pieces of precision='mixed': (1) unpool split=2 reads the hi halves of a pair tensor,
    (2) addend_pair_hi adds the hi halves of a pair tensor in a bf16 conv, (3) nothing else changes -- each equals
    the plain-bf16 call on the extracted hi tensor bit for bit."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(3)
    N, H, W, C = 2, 20, 26, 64
    x = torch.randn(N, H, W, C, device=cuda).to(torch.bfloat16)
    _, mask = K.maxpool2(x, True)
    u_pair = K.pack_nchw(torch.randn(N, C, H // 2, W // 2, device=cuda), C, split=True)      # [N,h,w,2C]
    u_hi = u_pair[..., :C].contiguous()
    win = (3, 2, 14, 21)
    assert torch.equal(K.unpool2(u_pair, mask, H, W, window=win, split=2), K.unpool2(u_hi, mask, H, W, window=win))
    Wt = (torch.randn(128, 9 * C, device=cuda) / (9 * C) ** 0.5).to(torch.bfloat16)
    b = torch.randn(128, device=cuda)
    add_pair = K.pack_nchw(torch.randn(N, 128, H + 4, W + 4, device=cuda), 128, split=True)
    add_hi = add_pair[..., :128].contiguous()
    for cout in (128, 64):                                   # CTA-pair kernel and halo-tile kernel epilogues
        got = K.conv2d(x, Wt[:cout].contiguous(), b[:cout].contiguous(), 3, 3, 1, relu=False,
                       addend=add_pair[..., :2 * cout].contiguous() if cout == 128 else torch.cat([add_pair[..., :64], add_pair[..., 128:192]], 3).contiguous(),
                       addend_off=(2, 1), addend_pair_hi=True)
        ref = K.conv2d(x, Wt[:cout].contiguous(), b[:cout].contiguous(), 3, 3, 1, relu=False,
                       addend=add_hi[..., :cout].contiguous(), addend_off=(2, 1))
        assert torch.equal(got, ref)


NPACK_CASES = [
    # N, H, W, window (oh0, ow0, OH, OW), pad
    (2, 40, 132, None, 1),                      # whole map, OW multiple of 4
    (1, 37, 45, None, 1),                       # odd sizes, OW % 4 != 0 (ragged last pixel group)
    (3, 64, 200, (1, 1, 62, 198), 1),           # the crop window of up_conv1 (origin 1, 1)
    (2, 50, 70, (5, 7, 24, 33), 1),             # odd window origin and extent
    (1, 362, 482, (1, 1, 360, 480), 1),         # BASELINE size: 4 x 120-pixel tiles per line
]


@pytest.mark.parametrize('case', NPACK_CASES, ids=[str(c) for c in NPACK_CASES])
def test_npack_logits_conv_matches_plain_kernel_and_reference(cuda, case):
    """The N-packed kernel of the 16-channel logits conv (four adjacent output pixels per accumulator row, TMA boxes with
    element stride 4): same result as the plain halo-tile kernel up to fp32 summation order, and as an fp32 convolution of
    the same bf16-rounded operands."""
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv, pack_npack16
    N, H, W, window, pad = case
    torch.manual_seed(5)
    x = torch.randn(N, 64, H, W, device=cuda).to(torch.bfloat16)
    Wt = (torch.randn(11, 64, 3, 3, device=cuda) / 24).to(torch.bfloat16).float()
    b = torch.randn(11, device=cuda)
    xs = x.permute(0, 2, 3, 1).contiguous()
    Wk, bk = pack_conv(Wt, b, [(64, 64)], 16, cuda)
    plain = K.conv2d(xs, Wk, bk, 3, 3, pad, relu=False, window=window, out_f32=True)
    assert K.last_conv_plan()[0] == 2
    packed = K.conv2d(xs, Wk, bk, 3, 3, pad, relu=False, window=window, out_f32=True, weight_npack=pack_npack16(Wk))
    assert K.last_conv_plan() == (3, 16, 64), K.last_conv_plan()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), Wt, b, padding=pad)
    if window:
        oh0, ow0, OH, OW = window
        ref = ref[:, :, oh0:oh0 + OH, ow0:ow0 + OW]
    got = packed[..., :11].permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    assert float((packed - plain).abs().max()) < 2e-5 * max(1.0, float(plain.abs().max()))
    assert float(packed[..., 11:].abs().max()) == 0.0

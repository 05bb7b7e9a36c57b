"""Out-of-bounds / stray-write guards (compute-sanitizer is closed on this pool, profiles/r02_sanitizer_unavailable.md):
every output tensor is a slice of a larger sentinel-filled allocation; after the launch the guard words around the slice
must be intact, and so must every element of the slice the kernel is documented not to write."""
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096        # elements before and after


def _guarded(shape, dtype, device, sentinel):
    n = 1
    for s in shape:
        n *= s
    flat = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device=device)
    return flat, flat[GUARD:GUARD + n].view(*shape)


def _intact(flat, sentinel):
    return bool((flat[:GUARD] == sentinel).all()) and bool((flat[-GUARD:] == sentinel).all())


CASES = [
    # N, H, W, C, Cout, pad, window, pool, split
    (2, 37, 45, 64, 128, 1, None, False, False),            # pair kernel
    (2, 37, 45, 64, 64, 1, (3, 5, 20, 31), False, False),   # halo kernel, window
    (2, 38, 46, 64, 64, 1, None, True, False),              # halo kernel, fused pool + mask
    (2, 37, 45, 64, 256, 1, (4, 6, 22, 30), True, False),   # pair kernel, fused pool on an even window
    (1, 30, 40, 16, 64, 100, None, True, True),             # halo split, 16-channel K blocks, pool
    (2, 33, 29, 64, 64, 1, None, False, True),              # halo split, plain
    (1, 17, 21, 128, 256, 1, None, True, True),             # pair kernel, split pool
    (2, 40, 56, 64, 16, 1, (5, 7, 24, 32), False, False),   # 16-channel logits conv (fp32 out)
]


@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_conv_writes_stay_inside_their_tensors(cuda, case):
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    N, H, W, C, Cout, pad, window, pool, split = case
    torch.manual_seed(0)
    x = K.pack_nchw(torch.randn(N, C, H, W, device=cuda), C, split=split)
    Wk, bk = pack_conv(torch.randn(Cout, C, 3, 3, device=cuda) / (9 * C) ** 0.5, torch.randn(Cout, device=cuda), [(C, C)], Cout, cuda, split=split)
    fOH, fOW = H + 2 * pad - 2, W + 2 * pad - 2
    oh0, ow0, OH, OW = window if window else (0, 0, fOH, fOW)
    cm = 2 if split else 1
    if pool:
        fp, pooled = _guarded((N, OH // 2, OW // 2, cm * Cout), torch.bfloat16, cuda, 7.0)
        fm, mask = _guarded((N, OH // 2, OW // 2, Cout // 8), torch.int32, cuda, 0x5A5A5A5A)
        K.conv2d(x, Wk, bk, 3, 3, pad, relu=True, window=window, pooled=pooled, pool_mask=mask, split=split)
        torch.cuda.synchronize()
        assert _intact(fp, 7.0) and _intact(fm, 0x5A5A5A5A)
        assert bool((mask != 0x5A5A5A5A).all())                       # every mask word was produced
    else:
        f32 = Cout == 16
        fo, out = _guarded((N, OH, OW, cm * Cout), torch.float32 if f32 else torch.bfloat16, cuda, 7.0)
        K.conv2d(x, Wk, bk, 3, 3, pad, relu=False, window=window, out=out, out_f32=f32, split=split)
        torch.cuda.synchronize()
        assert _intact(fo, 7.0)
        assert bool(torch.isfinite(out.float()).all())


def test_fused_update_and_streaming_kernels_stay_inside_their_tensors(cuda):
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200._packing import pack_conv
    torch.manual_seed(1)
    N, H, W, C = 2, 34, 50, 64
    x = K.pack_nchw(torch.randn(N, C, H, W, device=cuda), C)
    from iterative_inference_segm_b200._packing import pack_npack16
    Wk, bk = pack_conv(torch.randn(11, C, 3, 3, device=cuda) / 24, torch.randn(11, device=cuda), [(C, C)], 16, cuda)
    win = (1, 1, H - 2, W - 2)
    for ysp, npk in ((False, None), (True, None), (False, pack_npack16(Wk)), (True, pack_npack16(Wk))):
        fy, y = _guarded((N, 11, H - 2, W - 2), torch.float32, cuda, 0.25)
        fb, yb = _guarded((N, H - 2, W - 2, 32 if ysp else 16), torch.bfloat16, cuda, 7.0)
        acc = torch.zeros(N, dtype=torch.int64, device=cuda)
        K.conv2d(x, Wk, bk, 3, 3, 1, relu=False, window=win, out_f32=True, weight_npack=npk,
                 update=dict(y=y, y_bf16=yb, active=torch.ones(N, dtype=torch.int32, device=cuda), norm_acc=acc, step=0.05, C=11, y_split=ysp))
        torch.cuda.synchronize()
        assert _intact(fy, 0.25) and _intact(fb, 7.0) and bool((yb != 7.0).all()) and bool((acc > 0).all())
    # pool / unpool (windowed, all three split modes) / pack
    xs = torch.relu(torch.randn(N, H, W, C, device=cuda)).to(torch.bfloat16)
    fp, pooled = _guarded((N, H // 2, W // 2, C), torch.bfloat16, cuda, 7.0)
    fm, mask = _guarded((N, H // 2, W // 2, C // 8), torch.int32, cuda, 0x5A5A5A5A)
    K.maxpool2(xs, True, pooled=pooled, mask=mask)
    for split in (False, True, 2):
        cu = 2 * C if split else C
        co = C if split == 2 else cu
        u = torch.randn(N, H // 2, W // 2, cu, device=cuda).to(torch.bfloat16)
        fo, out = _guarded((N, 21, 30, co), torch.bfloat16, cuda, 7.0)
        K.unpool2(u, mask, H, W, out=out, window=(5, 7, 21, 30), split=split)
        torch.cuda.synchronize()
        assert _intact(fo, 7.0) and bool((out != 7.0).all())
    fk, packed = _guarded((N, H, W, 128), torch.bfloat16, cuda, 7.0)
    K.pack_nchw(torch.randn(N, C, H, W, device=cuda), C, out=packed, split=True)
    torch.cuda.synchronize()
    assert _intact(fp, 7.0) and _intact(fm, 0x5A5A5A5A) and _intact(fk, 7.0)

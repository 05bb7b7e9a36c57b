"""GPU parity of the bandwidth-bound kernels (pool+mask, unpool, softmax/update, metrics, layout)
against the CPU oracle.  Integer / mask / index work is bit-exact; float work states its tolerance."""
import numpy as np
import pytest
import torch

from oracle import lasagne_semantics as L, metrics as M

pytestmark = pytest.mark.gpu


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _mask_to_dense(mask, C):
    """[N,H2,W2,C/8] tie-mask words -> [N,C,2*H2,2*W2] 0/1; channel j of a group, position pos = 2*dy+dx,
    sits at bit 16*(j&1) + 4*(j>>1) + pos (include/iiseg.h)."""
    N, H2, W2, C8 = mask.shape
    m = mask.cpu().numpy().astype(np.uint32)
    out = np.zeros((N, C, 2 * H2, 2 * W2), np.float32)
    for j in range(8):
        for pos in range(4):
            bit = ((m >> (16 * (j & 1) + 4 * (j >> 1) + pos)) & 1).astype(np.float32)     # [N,H2,W2,C8]
            out[:, j::8, (pos >> 1)::2, (pos & 1)::2] = bit.transpose(0, 3, 1, 2)
    return out


@pytest.mark.parametrize('N,C,H,W', [(2, 64, 37, 45), (1, 128, 17, 21), (1, 8, 2, 2), (3, 256, 8, 10), (1, 64, 139, 169)])
def test_pool_mask_unpool_bit_exact(cuda, N, C, H, W):
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(0)
    # post-ReLU-like data with many exact ties (zeros, repeated values)
    x = torch.relu(torch.randn(N, C, H, W)).mul(4).round().div(4).to(torch.bfloat16)
    xd = _nhwc(x).to(cuda)
    pooled, mask = K.maxpool2(xd, with_mask=True)
    ref_pool = L.maxpool2(x.float())
    assert torch.equal(pooled.cpu().float().permute(0, 3, 1, 2), ref_pool)
    ref_mask = L.tie_mask(x.float())[:, :, :2 * (H // 2), :2 * (W // 2)]
    assert np.array_equal(_mask_to_dense(mask, C), ref_mask.numpy())
    u = torch.randn(N, C, H // 2, W // 2).to(torch.bfloat16)
    out = K.unpool2(_nhwc(u).to(cuda), mask, H, W)
    ref = L.depool2d(u.float(), x.float())
    assert torch.equal(out.cpu().float().permute(0, 3, 1, 2), ref)      # includes the zero odd row/col


def test_unpool_window_bit_exact(cuda):
    """Windowed DePool2D: an odd-origin output window fed from a window of the pooled map equals the
    same slice of the full unpool (incl. the zero trailing odd row/col)."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(5)
    N, C, H, W = 2, 64, 23, 31
    x = torch.relu(torch.randn(N, C, H, W)).mul(2).round().div(2).to(torch.bfloat16)
    _, mask = K.maxpool2(_nhwc(x).to(cuda), with_mask=True)
    u = torch.randn(N, C, H // 2, W // 2).to(torch.bfloat16)
    full = L.depool2d(u.float(), x.float())
    for (h0, w0, OH, OW) in [(3, 5, 17, 22), (0, 0, 23, 31), (7, 8, 16, 23), (22, 30, 1, 1)]:
        ph0, pw0 = h0 // 2, w0 // 2
        ph1, pw1 = min((h0 + OH - 1) // 2 + 1, H // 2), min((w0 + OW - 1) // 2 + 1, W // 2)
        ph0, pw0 = min(ph0, H // 2 - 1), min(pw0, W // 2 - 1)      # trailing odd row/col: any non-empty u window
        ph1, pw1 = max(ph1, ph0 + 1), max(pw1, pw0 + 1)
        uw = _nhwc(u)[:, ph0:ph1, pw0:pw1].contiguous().to(cuda)
        out = K.unpool2(uw, mask, H, W, u_origin=(ph0, pw0), window=(h0, w0, OH, OW))
        assert torch.equal(out.cpu().float().permute(0, 3, 1, 2), full[:, :, h0:h0 + OH, w0:w0 + OW])


def test_pool_without_mask(cuda):
    from iterative_inference_segm_b200 import _kernels as K
    x = torch.randn(2, 64, 10, 14).to(torch.bfloat16)
    pooled = K.maxpool2(_nhwc(x).to(cuda), with_mask=False)
    assert torch.equal(pooled.cpu().float().permute(0, 3, 1, 2), L.maxpool2(x.float()))


def test_pack_unpack_roundtrip(cuda):
    from iterative_inference_segm_b200 import _kernels as K
    x = torch.randn(2, 11, 9, 13)
    p = K.pack_nchw(x.to(cuda), 64)
    assert p.shape == (2, 9, 13, 64)
    assert torch.equal(p.cpu()[..., :11].float().permute(0, 3, 1, 2), x.to(torch.bfloat16).float())
    assert float(p.cpu()[..., 11:].float().abs().max()) == 0.0
    back = K.unpack_nhwc(p, 11)
    assert torch.equal(back.cpu(), x.to(torch.bfloat16).float())


def test_softmax_update_matches_loop_body(cuda):
    """iterative_inference.py:267-277 on synthetic logits: tolerance 2e-6 (expf vs torch.exp)."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(2)
    N, C, H, W = 3, 11, 19, 23
    logits = torch.zeros(N, H, W, 16)
    logits[..., :C] = torch.randn(N, H, W, C) * 3
    y = torch.softmax(torch.randn(N, C, H, W), 1)
    p = torch.softmax(logits[..., :C].permute(0, 3, 1, 2), 1)
    g = y - p
    step = 0.3
    y_ref = torch.clamp(y - step * g, 0, 1)
    norm_ref = torch.linalg.vector_norm(g, dim=1).mean((1, 2))
    yd = y.to(cuda).clone()
    yb = torch.zeros(N, H, W, 64, dtype=torch.bfloat16, device=cuda)
    pout = torch.zeros(N, C, H, W, device=cuda)
    active = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
    partial = torch.zeros(N, K.update_blocks(H, W), device=cuda)
    norm = torch.zeros(N, device=cuda)
    n_exec = torch.zeros(N, dtype=torch.int32, device=cuda)
    K.softmax_update(logits.to(cuda), yd, yb, active, partial, step, p_out=pout)
    K.norm_finalize(partial, norm, active, n_exec, H, W, 1e-3)
    yd, pout, norm, yb = yd.cpu(), pout.cpu(), norm.cpu(), yb.cpu()
    for n in (0, 2):
        assert float((yd[n] - y_ref[n]).abs().max()) < 2e-6
        assert float((pout[n] - p[n]).abs().max()) < 2e-6
        assert abs(float(norm[n] - norm_ref[n])) < 1e-6
        assert torch.equal(yb[n, :, :, :C].float(), yd[n].permute(1, 2, 0).to(torch.bfloat16).float())
    assert torch.equal(yd[1], y[1])                    # frozen image untouched
    assert n_exec.cpu().tolist() == [1, 0, 1] and active.cpu().tolist() == [1, 0, 1]
    # eps above the norm: the image is switched off AFTER its update
    K.norm_finalize(partial, torch.zeros(N, device=cuda), active, n_exec, H, W, 1e9)
    assert active.cpu().tolist() == [0, 0, 0] and n_exec.cpu().tolist() == [2, 0, 2]
    # de_fn
    grad = torch.zeros(N, C, H, W, device=cuda)
    K.softmax_grad(logits.to(cuda), y.to(cuda), grad)
    assert float((grad.cpu() - g).abs().max()) < 2e-6


@pytest.mark.parametrize('H,W', [(2, 2), (37, 45), (360, 480)])
def test_metrics_bit_exact(cuda, H, W):
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200.functions import function_val, jaccard_from_cm
    torch.manual_seed(3)
    N, C = 2, 11
    y = torch.softmax(torch.randn(N, C, H, W) * 2, 1)
    y[0, :, 0, 0] = 1.0 / C                               # exact argmax tie -> first index
    lab = torch.randint(0, C + 1, (N, H, W))
    onehot = torch.nn.functional.one_hot(lab, C + 1).permute(0, 3, 1, 2).float().contiguous()
    cm_ref = M.confusion_matrix(y.numpy(), onehot.numpy(), C)
    corr, valid = M.accuracy_counts(y.numpy(), onehot.numpy(), [C])
    for kw in ({'onehot': onehot.to(cuda)}, {'labels': lab.to(torch.int32).to(cuda)}):
        cm = torch.zeros(N, C * C, dtype=torch.int64, device=cuda)
        cnt = torch.zeros(N, 2, dtype=torch.int64, device=cuda)
        se = torch.zeros(N, 2, dtype=torch.float64, device=cuda)
        K.metrics_accumulate(y.to(cuda), cm, cnt, se, void_label=C, **kw)
        assert np.array_equal(cm.sum(0).cpu().numpy().reshape(C, C), cm_ref)          # bit-exact
        assert cnt.sum(0).cpu().tolist() == [corr, valid]
        mse = float(se[:, 0].sum() / se[:, 1].sum())
        assert abs(mse - float(M.squared_error(y.numpy(), onehot.numpy(), C))) < 1e-6
    acc, jacc, mse = function_val(C, [C])(y.numpy(), onehot.numpy())
    assert jacc.dtype == np.float32 and np.array_equal(jacc, M.jaccard(y.numpy(), onehot.numpy(), C))
    assert acc == M.accuracy(y.numpy(), onehot.numpy(), [C])
    assert np.array_equal(jaccard_from_cm(cm_ref), M.jaccard_from_cm(cm_ref))


def test_metrics_active_gating_and_accumulation(cuda):
    from iterative_inference_segm_b200 import _kernels as K
    N, C, H, W = 3, 11, 8, 8
    y = torch.softmax(torch.randn(N, C, H, W), 1).to(cuda)
    lab = torch.randint(0, C + 1, (N, H, W), dtype=torch.int32).to(cuda)
    cm = torch.zeros(N, C * C, dtype=torch.int64, device=cuda)
    cnt = torch.zeros(N, 2, dtype=torch.int64, device=cuda)
    se = torch.zeros(N, 2, dtype=torch.float64, device=cuda)
    active = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
    K.metrics_accumulate(y, cm, cnt, se, labels=lab, active=active, void_label=C)
    K.metrics_accumulate(y, cm, cnt, se, labels=lab, active=active, void_label=C)
    assert int(cm[1].sum()) == 0 and int(cnt[1].sum()) == 0
    assert int(cm[0].sum()) == 2 * int((lab[0] < C).sum())


@pytest.mark.parametrize('split', [False, True])
def test_widen_nhwc_bf16_to_f32(cuda, split):
    """iiseg_widen_nhwc_bf16_to_f32: NHWC bf16 (or the (hi | lo) pair) -> NHWC fp32 = hi (+ lo), exactly."""
    from iterative_inference_segm_b200 import _kernels as K
    torch.manual_seed(3)
    x = torch.randn(3, 24, 5, 7, device=cuda)
    src = K.pack_nchw(x, 24, split=split)                       # [3,5,7,24] or [3,5,7,48]
    out = torch.full((3, 5, 7, 24), -1.0, device=cuda)
    K.widen_nhwc(src, out, split=split)
    want = src[..., :24].float() + (src[..., 24:].float() if split else 0.0)
    assert torch.equal(out, want)
    assert float((out.permute(0, 3, 1, 2) - x).abs().max()) < (2e-5 if split else 2e-2)
    part = torch.full((3, 5, 7, 24), -1.0, device=cuda)         # a batch slice, as the bn=1 mask pass calls it
    K.widen_nhwc(src[1:2], part[1:2], split=split)
    assert torch.equal(part[1], want[1]) and float(part[0].max()) == -1.0 and float(part[2].max()) == -1.0

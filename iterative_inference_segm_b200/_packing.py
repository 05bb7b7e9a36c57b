"""Repacks Lasagne-layout parameters into the kernels' layouts (done once at build
time, on the device).

Lasagne conv W is (out, in, kh, kw) (flip_filters=False -> cross-correlation);
the GEMM B operand is bf16 [Cout_pad][kh*kw][Cin_pad] (K-major), where the input
channels may be the concatenation of several sources, each padded to a multiple
of 64 on its own.  Deconv W is (in, out, kh, kw).
"""
import numpy as np
import torch

from ._kernels import pad_channels


def _as_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)


def pack_conv(W, b, splits, cout_pad, device, split=False):
    """W (Cout, sum(real), R, S), splits = [(real, padded), ...] per concatenated source.
    Returns (bf16 [cout_pad, R*S*sum(padded)], fp32 bias [cout_pad]).
    split=True (fp32-accurate variant): K per tap is (W_hi | W_hi | W_lo) with W_hi = bf16(W),
    W_lo = bf16(W - W_hi), matching the (hi | lo | hi) activation views of _kernels.conv2d."""
    W = _as_f32(W, device)
    b = _as_f32(b, device)
    Cout, Cin, R, S = W.shape
    assert Cin == sum(r for r, _ in splits), (Cin, splits)
    parts, c0 = [], 0
    for real, padded in splits:
        blk = torch.zeros((cout_pad, R, S, padded), dtype=torch.float32, device=device)
        blk[:Cout, :, :, :real] = W[:, c0:c0 + real].permute(0, 2, 3, 1)
        parts.append(blk)
        c0 += real
    Wf = torch.cat(parts, dim=3)
    if split:
        hi = Wf.to(torch.bfloat16)
        lo = (Wf - hi.float()).to(torch.bfloat16)
        Wk = torch.cat([hi, hi, lo], dim=3).reshape(cout_pad, -1).contiguous()
    else:
        Wk = Wf.reshape(cout_pad, -1).to(torch.bfloat16).contiguous()
    bk = torch.zeros((cout_pad,), dtype=torch.float32, device=device)
    bk[:Cout] = b
    return Wk, bk


def pack_npack16(Wk):
    """Re-pack a 16-row 3x3 filter bank of one 64-channel source, bf16 [16, 9*64] (pack_conv's layout: tap-major, then
    channel), for the N-packed kernel (iiseg_conv_desc.weight_npack): for filter column s and input-line phase c = j + r
    (0..5), the 16-row blocks W[.][c - j][s][.] of the output pixels j = max(0, c-2) .. min(3, c) one after the other, in
    (s, c) order; three zero blocks follow block (0, 0) (that instruction initialises all four pixel groups).
    Returns bf16 [39*16, 64]."""
    assert Wk.dtype == torch.bfloat16 and tuple(Wk.shape) == (16, 9 * 64), tuple(Wk.shape)
    W = Wk.view(16, 3, 3, 64)
    blocks = []
    for s_ in range(3):
        for c in range(6):
            for j in range(max(0, c - 2), min(3, c) + 1):
                blocks.append(W[:, c - j, s_, :])
            if s_ == 0 and c == 0:
                blocks += [torch.zeros_like(W[:, 0, 0, :])] * 3
    out = torch.cat(blocks, dim=0).contiguous()
    assert tuple(out.shape) == (39 * 16, 64)
    return out


def pack_deconv16(W, b, device, scale=1.0):
    """Lasagne Deconv2DLayer W (in, out, k, k) -> Wt[a][b][ci][co] = W[ci, co, k-1-a, k-1-b]
    (flip: the layer is the input-gradient of a true convolution), zero-padded to 16x16."""
    W = _as_f32(W, device) * scale
    b = _as_f32(b, device) * scale
    Cin, Cout, k, _ = W.shape
    assert Cin <= 16 and Cout <= 16
    Wt = torch.zeros((k, k, 16, 16), dtype=torch.float32, device=device)
    Wt[:, :, :Cin, :Cout] = W.flip(2, 3).permute(2, 3, 0, 1)
    bk = torch.zeros((16,), dtype=torch.float32, device=device)
    bk[:Cout] = b
    return Wt.contiguous(), bk


def load_npz_params(path):
    """Positional checkpoint arr_0..arr_k (models/DAE_h.py:52-57, models/fcn8.py:177-180)."""
    with np.load(path) as f:
        return [f['arr_%d' % i] for i in range(len(f.files))]


__all__ = ['pack_conv', 'pack_deconv16', 'pack_npack16', 'load_npz_params', 'pad_channels']

"""Which conv layers of DAE_h need fp32-grade arithmetic?  (test infrastructure, CPU only)

Emulates, inside the oracle's DAE forward pass, the arithmetic variants of the CUDA conv kernel
layer by layer and measures the parity numbers of BASELINE.json (max-abs on p / y per iteration,
argmax agreement) against the fp32 oracle, next to the oracle's own fp32-vs-fp64 noise floor:

    'f32'  : the oracle's conv (what 'x3' approximates to ~2^-17)
    'x3'   : operands as (hi, lo) bf16 pairs, hi*hi + lo*hi + hi*lo, output stored as a pair
    'bf16' : operands rounded to bf16, fp32 accumulation, output stored as bf16
    'f16'  : operands rounded to IEEE half, fp32 accumulation, output stored as half

    python -m oracle.precision_mix [H W N] -- prints one table per candidate mix

The result decides the cheapest mix the parity-grade CUDA variant may use.
"""
import sys
import time

import torch
import torch.nn.functional as F

from . import lasagne_semantics as L
from . import nets, weights


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def hf(x):
    return x.to(torch.float16).to(torch.float32)


def pair(x):
    hi = bf(x)
    return hi, bf(x - hi)


def conv_mode(x, W, b, pad, mode):
    if pad == 'same':
        pad = W.shape[2] // 2
    if mode == 'f32':
        return F.conv2d(x, W, b, padding=pad)
    if mode == 'bf16':
        return F.conv2d(bf(x), bf(W), None, padding=pad) + b.view(1, -1, 1, 1)
    if mode == 'f16':       # operands rounded to IEEE half (11 significant bits), fp32 accumulation
        return F.conv2d(hf(x), hf(W), None, padding=pad) + b.view(1, -1, 1, 1)
    if mode == 'x3':
        xh, xl = pair(x)
        wh, wl = pair(W)
        return (F.conv2d(xh, wh, None, padding=pad) + F.conv2d(xl, wh, None, padding=pad)
                + F.conv2d(xh, wl, None, padding=pad)) + b.view(1, -1, 1, 1)
    if mode == 'x2a':       # activations as a pair, weights rounded: hi*Whi + lo*Whi
        xh, xl = pair(x)
        return F.conv2d(xh + xl, bf(W), None, padding=pad) + b.view(1, -1, 1, 1)
    if mode == 'x2w':       # activations rounded, weights as a pair: hi*Whi + hi*Wlo
        wh, wl = pair(W)
        return F.conv2d(bf(x), wh + wl, None, padding=pad) + b.view(1, -1, 1, 1)
    raise ValueError(mode)


def store(x, mode):
    """What the kernel writes: bf16, the (hi, lo) pair (16 significant bits), or fp32."""
    if mode == 'bf16':
        return bf(x)
    if mode == 'f16':
        return hf(x)
    if mode in ('x3', 'x2a'):
        hi, lo = pair(x)
        return hi + lo
    if mode == 'x2w':
        return bf(x)
    return x


def dae_forward_mix(params, y, h, padding, down_modes, up_modes, n_pool=4, total=6):
    """oracle.nets.dae_forward with a per-layer arithmetic mode (models/fcn_down.py:77-136,
    models/fcn_up.py:65-113)."""
    Wd = [(params[2 * i], params[2 * i + 1]) for i in range(total)]
    Wu = [(params[2 * (total + i)], params[2 * (total + i) + 1]) for i in range(total)]
    x = store(y, down_modes[0])
    pre, pools = [], []
    for p in range(total):
        m = down_modes[p]
        x = torch.relu(conv_mode(x, *Wd[p], pad=padding if p == 0 else 'same', mode=m))
        x = store(x, m)
        pre.append(x)
        x = L.maxpool2(x)
        pools.append(x)
        if p + 1 == n_pool:
            x = torch.cat([store(h, m), x], dim=1)
    u = pools[-1]
    for i, p in enumerate(range(total, 0, -1)):
        m = up_modes[i]
        u = L.depool2d(store(u, m), pre[p - 1])
        u = conv_mode(u, *Wu[i], pad='same', mode=m)
        if p > 1:
            a, b = L.center_crop_pair(u, pools[p - 2])
            u = store(a + (bf(b) if m == 'bf16' else hf(b) if m == 'f16' else b), m)      # bf16 expanding path: skip-sum with the hi half of the pool pair
        else:
            u = L.center_crop_to(u, y.shape[2], y.shape[3])
    return L.channel_softmax(u)


MIXES = {
    'all-x3': (['x3'] * 6, ['x3'] * 6),
    'down-x3/up-bf16': (['x3'] * 6, ['bf16'] * 6),
    'down1-4-x3/rest-bf16': (['x3'] * 4 + ['bf16'] * 2, ['bf16'] * 6),
    'down1-5-x3/rest-bf16': (['x3'] * 5 + ['bf16'] * 1, ['bf16'] * 6),
    'all-bf16': (['bf16'] * 6, ['bf16'] * 6),
    'all-f16': (['f16'] * 6, ['f16'] * 6),
    'down-x3/up-f16': (['x3'] * 6, ['f16'] * 6),
    'down1-2-f16,3-6-x3/up-bf16': (['f16'] * 2 + ['x3'] * 4, ['bf16'] * 6),
    'down1-4-x3,5-6-f16/up-bf16': (['x3'] * 4 + ['f16'] * 2, ['bf16'] * 6),
    'down1-f16,2-6-x3/up-bf16': (['f16'] + ['x3'] * 5, ['bf16'] * 6),
    'down1-5-x3,6-f16/up-bf16': (['x3'] * 5 + ['f16'], ['bf16'] * 6),
    'down-f16/up-bf16': (['f16'] * 6, ['bf16'] * 6),
    'down5-6-x3/rest-bf16': (['bf16'] * 4 + ['x3'] * 2, ['bf16'] * 6),
    'down2-6-x3/rest-bf16': (['bf16'] + ['x3'] * 5, ['bf16'] * 6),
    'down3-6-x3/rest-bf16': (['bf16'] * 2 + ['x3'] * 4, ['bf16'] * 6),
    'down-x2a/up-bf16': (['x2a'] * 6, ['bf16'] * 6),
    'down1-4-x2a,5-6-x3/up-bf16': (['x2a'] * 4 + ['x3'] * 2, ['bf16'] * 6),
}


def main(H=224, W=224, N=12, step=0.05, which=None):
    torch.manual_seed(0)
    X, _, _ = weights.synthetic_batch(1, H, W)
    pf = weights.synthetic_fcn8_params(3, 11, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(11, 512, seed=1, out_gain=0.1)
    h, y0 = nets.fcn8_forward(pf, X, 11)
    pd64 = [p.double() for p in pd]
    names = which or list(MIXES)
    ys = {n: y0.clone() for n in names}
    y32, y64 = y0.clone(), y0.double()
    print('%dx%d, %d iterations; columns: teacher-forced p max-abs | free-running y max-abs | argmax disagreement' % (H, W, N))
    for it in range(N):
        t = time.time()
        p32 = nets.dae_forward(pd, y32, h, 100)
        p64 = nets.dae_forward(pd64, y64, h.double(), 100)
        line = 'it %2d  fp64-noise p %.1e y %.1e |' % (it + 1, float((p32.double() - p64).abs().max()), float((y32.double() - y64).abs().max()))
        for n in names:
            dm, um = MIXES[n]
            p_tf = dae_forward_mix(pd, y32, h, 100, dm, um)            # teacher-forced: the oracle's y in
            p_fr = dae_forward_mix(pd, ys[n], h, 100, dm, um)
            ys[n] = torch.clamp(ys[n] - step * (ys[n] - p_fr), 0, 1)
            y_next = torch.clamp(y32 - step * (y32 - p32), 0, 1)
            line += ' %s: %.1e %.1e %.4f%% |' % (n, float((p_tf - p32).abs().max()), float((ys[n] - y_next).abs().max()),
                                                 100 * float((ys[n].argmax(1) != y_next.argmax(1)).float().mean()))
        y32 = torch.clamp(y32 - step * (y32 - p32), 0, 1)
        y64 = torch.clamp(y64 - step * (y64 - p64), 0, 1)
        print(line + ' (%.0fs)' % (time.time() - t), flush=True)


if __name__ == '__main__':
    a = sys.argv[1:]
    nums = [int(v) for v in a[:3]]
    main(*nums, which=a[3:] or None)

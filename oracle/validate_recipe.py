"""Validate the synthetic weight recipe in the oracle:
free-running fp32-vs-fp64 drift must stay << 2e-3, argmax margins must not be
degenerate, and no image may early-exit.  Prints one line per iteration.

    python -m oracle.validate_recipe [H W N gain out_gain]
"""
import sys
import time

import torch

from . import nets, weights, loop


def main(H=224, W=224, N=50, gain=10.0, out_gain=1.0, step=0.05):
    X, L, _ = weights.synthetic_batch(1, H, W)
    pf = weights.synthetic_fcn8_params(3, 11, seed=0, logit_gain=gain)
    pd = weights.synthetic_dae_params(11, 512, seed=1, out_gain=out_gain)
    t = time.time()
    h, y0 = nets.fcn8_forward(pf, X, 11)
    print('fcn8 %.1fs y0 range [%.4f, %.4f] top1 mean %.3f' % (
        time.time() - t, float(y0.min()), float(y0.max()), float(y0.max(1)[0].mean())))
    pd64 = [p.double() for p in pd]
    y32, y64 = y0.clone(), y0.double()
    h64 = h.double()
    for it in range(N):
        t = time.time()
        p32 = nets.dae_forward(pd, y32, h, 100)
        g32 = y32 - p32
        y32 = torch.clamp(y32 - step * g32, 0, 1)
        p64 = nets.dae_forward(pd64, y64, h64, 100)
        g64 = y64 - p64
        y64 = torch.clamp(y64 - step * g64, 0, 1)
        norm = float(torch.linalg.vector_norm(g32, dim=1).mean())
        drift = float((y32.double() - y64).abs().max())
        pdrift = float((p32.double() - p64).abs().max())
        agree = float((y32.argmax(1) == y64.argmax(1)).float().mean())
        top2 = torch.topk(y64, 2, dim=1)[0]
        margin = (top2[:, 0] - top2[:, 1])
        print('it %2d norm %.5f  y drift %.2e  p drift %.2e  argmax agree %.5f  '
              'margin med %.4f p1%% %.2e  (%.1fs)' % (
                  it + 1, norm, drift, pdrift, agree, float(margin.median()),
                  float(margin.flatten().kthvalue(max(1, margin.numel() // 100))[0]),
                  time.time() - t), flush=True)


if __name__ == '__main__':
    a = sys.argv[1:]
    main(*(int(a[i]) if i < 3 else float(a[i]) for i in range(len(a))))

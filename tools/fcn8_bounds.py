"""Per-launch times of one FCN8 forward pass (batch 10, 360x480), CUDA-event timed:
    python tools/fcn8_bounds.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(B=10, H=360, W=480, precision='bf16'):
    from iterative_inference_segm_b200 import synthetic
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.profiling import KernelTimer
    pf = synthetic.synthetic_fcn8_params(3, 11, seed=0, logit_gain=10.0)
    fcn = buildFCN8(3, None, n_classes=11, layer=['pool4', 'probs_dimshuffle'], params=pf, precision=precision)
    net = fcn[0].net
    X = torch.rand(B, 3, H, W, device='cuda')
    net.forward(X)
    timer = KernelTimer()
    with timer.recording():
        for _ in range(3):
            net.forward(X)
    tot = 0.0
    for (name, tag), v in timer.summary().items():
        ms = sum(v[1:]) / len(v[1:])
        tot += ms
        print('%-14s %-90s %8.1f us' % (name, str(tag)[:90], ms * 1e3))
    print('total %.1f us [precision=%s]' % (tot * 1e3, precision))


if __name__ == '__main__':
    main(precision=sys.argv[1] if len(sys.argv) > 1 else 'bf16')

"""Times the 12 conv launches of one DAE application (batch 10, 360x480) in isolation with CUDA
events; used with IISEG_CONV_DBG / IISEG_CONV_STAGES to find what bounds each layer."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(B=10, reps=10):
    from iterative_inference_segm_b200 import _kernels as K
    dev = 'cuda'
    sizes = [(558, 678), (279, 339), (139, 169), (69, 84), (34, 42), (17, 21)]
    down = [(64, 0, 64, 11), (64, 0, 128, 64), (128, 0, 256, 128), (256, 0, 512, 256), (512, 512, 1024, 1024), (1024, 0, 2048, 1024)]
    up = [(2048, 1024, 2048), (1024, 512, 1024), (512, 256, 512), (256, 128, 256), (128, 64, 128), (64, 16, 64)]
    layers = []
    for i, ((h, w), (c0, c1, co, creal)) in enumerate(zip(sizes, down)):
        hin, win = (360, 480) if i == 0 else (h, w)
        layers.append(('conv%d_1' % (i + 1), hin, win, c0, c1, co, 100 if i == 0 else 1, None, False, 2.0 * h * w * creal * co * 9))
    for i, ((h, w), (ci, co, creal)) in enumerate(zip(sizes[::-1], up)):
        p = 6 - i
        win_ = ((h - 360) // 2, (w - 480) // 2, 360, 480) if p == 1 else None
        hh, ww = (360, 480) if p == 1 else (h, w)
        layers.append(('up_conv%d' % p, h, w, ci, 0, co, 1, win_, p > 1, 2.0 * hh * ww * creal * (11 if p == 1 else co) * 9))
    tot = 0.0
    for name, h, w, c0, c1, co, pad, win_, addend, fl in layers:
        x0 = torch.randn(B, h, w, c0, device=dev).to(torch.bfloat16)
        x1 = torch.randn(B, h, w, c1, device=dev).to(torch.bfloat16) if c1 else None
        Wk = (torch.randn(co, 9 * (c0 + c1), device=dev) * 0.02).to(torch.bfloat16)
        b = torch.zeros(co, device=dev)
        oh, ow = (win_[2], win_[3]) if win_ else (h + 2 * pad - 2, w + 2 * pad - 2)
        out = torch.empty(B, oh, ow, co, dtype=torch.float32 if co == 16 else torch.bfloat16, device=dev)
        add = torch.randn(B, oh, ow, co, device=dev).to(torch.bfloat16) if addend else None
        for _ in range(2):
            K.conv2d(x0, Wk, b, 3, 3, pad, relu=not addend, src1=x1, addend=add, window=win_, out=out, out_f32=(co == 16))
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            K.conv2d(x0, Wk, b, 3, 3, pad, relu=not addend, src1=x1, addend=add, window=win_, out=out, out_f32=(co == 16))
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / reps * 1e3
        tot += us
        print('%-9s %4dx%-4d C %4d+%-4d -> %4d  %8.1f us  %7.1f TFLOP/s (useful)' % (name, h, w, c0, c1, co, us, fl * B / us / 1e6), flush=True)
        del x0, x1, Wk, out, add
    print('total %.1f us  [dbg=%s stages=%s]' % (tot, os.environ.get('IISEG_CONV_DBG'), os.environ.get('IISEG_CONV_STAGES')))


if __name__ == '__main__':
    main()

"""bn_relu_pack / channel_stats bandwidth at FC-DenseNet103 stack shapes (batch 10): algorithmic bytes / device time."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import _kernels as K

def ev(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B = 10
for (H, W, Cs, C) in [(360, 480, 112, 48), (360, 480, 112, 96), (360, 480, 256, 192), (360, 480, 256, 240), (180, 240, 384, 304),
                      (90, 120, 576, 464), (45, 60, 816, 656), (22, 30, 1088, 896)]:
    st = torch.randn(B, H, W, Cs, device='cuda')
    cp = (C + 63) // 64 * 64
    out = torch.empty(B, H, W, cp, dtype=torch.bfloat16, device='cuda')
    mean, istd = torch.zeros(Cs, device='cuda'), torch.ones(Cs, device='cuda')
    gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    ms = ev(lambda: K.bn_relu_pack(st, C, out, stats=(mean, istd), gamma=gamma, beta=beta, relu=True))
    nb = B * H * W * (C * 4 + cp * 2)
    print('bn_relu_pack %3dx%3d Cs %4d C %3d -> %3d: %7.1f us %6.0f GB/s' % (H, W, Cs, C, cp, ms * 1e3, nb / ms / 1e6))

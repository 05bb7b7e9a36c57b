"""Probe: K.wgrad_gemm against a torch matmul, for aligned and unaligned tap offsets."""
import sys
import torch
sys.path.insert(0, '.')
from iterative_inference_segm_b200 import _kernels as K

dev = torch.device('cuda:0')
torch.manual_seed(0)
for (M, cin, Kt, slabs, koffs) in [(16, 64, 4096, 1, [0, 8, 16]), (16, 64, 4096, 1, [0, 72, 144]), (128, 256, 8192, 4, [0, 72, 144]), (64, 16, 8192, 2, [0, 32, 64]),
                                   (256, 512, 4096, 2, [0, 32, 64])]:
    gT = torch.randn(M, Kt, device=dev).to(torch.bfloat16)
    xT = torch.randn(3 * cin, Kt, device=dev).to(torch.bfloat16)
    groups = [(s_ * cin, k) for k in koffs for s_ in range(3)]
    ld = (9 * cin + 1 + 63) // 64 * 64
    try:
        G = K.wgrad_gemm(gT, xT, cin, groups, slabs, ld)
        torch.cuda.synchronize()
    except Exception as e:
        print('FAILED', (M, cin, Kt, slabs, koffs), e)
        break
    xp = torch.cat([xT.float(), torch.zeros(3 * cin, 256, device=dev)], 1)
    ref = torch.cat([gT.float() @ xp[r:r + cin, k:k + Kt].t() for (r, k) in groups], 1)
    err = (G[:, :ref.shape[1]] - ref).abs().max().item()
    print((M, cin, Kt, slabs), 'koffs', koffs[:4], 'max err %.3e (ref max %.1f)' % (err, ref.abs().max().item()))

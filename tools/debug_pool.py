import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import _kernels as K
N, H, W, C, Cout, pad = 2, 61, 47, 128, 64, 1
torch.manual_seed(4)
cuda = 'cuda'
x = torch.randn(N, H, W, C, device=cuda).to(torch.bfloat16)
Wk = (torch.randn(Cout, 9 * C, device=cuda) / (9 * C) ** 0.5).to(torch.bfloat16)
b = torch.randn(Cout, device=cuda)
full = K.conv2d(x, Wk, b, 3, 3, pad, relu=True)
ref_p, ref_m = K.maxpool2(full, with_mask=True)
OH, OW = full.shape[1], full.shape[2]
pooled = torch.zeros((N, OH // 2, OW // 2, Cout), dtype=torch.bfloat16, device=cuda)
mask = torch.zeros((N, OH // 2, OW // 2, Cout // 8), dtype=torch.int32, device=cuda)
K.conv2d(x, Wk, b, 3, 3, pad, relu=True, pooled=pooled, pool_mask=mask)
bad = (pooled != ref_p)
print('pooled mismatches', int(bad.sum()), 'of', bad.numel())
idx = bad.nonzero()
print(idx[:20].tolist())
print('by n', bad.sum((1, 2, 3)).tolist()); print('by row', bad.sum((0, 2, 3)).tolist()); print('by col', bad.sum((0, 1, 3)).tolist()); print('by ch', bad.sum((0, 1, 2)).tolist())
print('mask mismatches', int((mask != ref_m).sum()))

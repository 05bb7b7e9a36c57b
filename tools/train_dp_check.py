"""Data-parallel DAE train step on N GPUs (torchrun): every rank trains on its shard of one global batch with the
global-denominator protocol + NCCL gradient sums; rank 0 repeats the step alone on the whole batch and compares
the updated weights (they agree up to bf16 / summation-order noise)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S, _kernels as K
from iterative_inference_segm_b200.sharding import World, shard_range
from iterative_inference_segm_b200.train_dae import DAETrainer

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
NCLS, B, H, W = 11, 4, 64, 80
pd = S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
X, L, _ = S.synthetic_batch(B, H, W, NCLS, seed=5)
L = L.cuda(); y = L[:, :NCLS].contiguous()
gen = torch.Generator(device='cuda').manual_seed(3)
hs = (((H + 198) // 2 // 2 // 2) // 2, ((W + 198) // 2 // 2 // 2) // 2)
h = K.pack_nchw(torch.relu(torch.randn((B, 512) + hs, device='cuda', generator=gen)), 512)
nm = torch.randn(y.shape, device='cuda', generator=gen); nk = torch.randn(y.shape, device='cuda', generator=gen)
lo, hi = shard_range(B, rank, world)
tr = DAETrainer(NCLS, 512, 100, pd, learning_rate=1e-3, noise=0.5)
tr.step(h[lo:hi].contiguous(), y[lo:hi].contiguous(), L[lo:hi].contiguous(), nm[lo:hi].contiguous(), nk[lo:hi].contiguous(), world=World())
torch.cuda.synchronize()
loss_dp = tr.loss_value()
if rank == 0:
    ref = DAETrainer(NCLS, 512, 100, pd, learning_rate=1e-3, noise=0.5)
    ref.step(h, y, L, nm, nk)
    worst = 0.0
    for a, b, p0 in zip(tr.params(), ref.params(), pd):
        step_ref = (b.cpu() - p0).norm()
        worst = max(worst, float((a - b).norm() / step_ref.clamp(min=1e-12)))
    print('dp%d vs single device: loss %.6f vs %.6f; worst relative difference of the weight update %.3e' % (world, loss_dp, ref.loss_value(), worst))
    assert abs(loss_dp - ref.loss_value()) < 1e-6 * abs(ref.loss_value()) + 1e-9 and worst < 0.05
dist.barrier()
dist.destroy_process_group()

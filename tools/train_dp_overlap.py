"""Data-parallel train step (torchrun, NCCL): step time with the gradient buckets all-reduced under backward, as a function of
the SMs left free for NCCL's CTAs (IISEG_DP_RESERVED_SMS) -- and against one blocking all-reduce / no exchange at all."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S, _kernels as K
from iterative_inference_segm_b200.sharding import World
from iterative_inference_segm_b200.train_dae import DAETrainer

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
NCLS, B, H, W = 11, 10, 224, 224
tr = DAETrainer(NCLS, 512, 100, S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1), learning_rate=1e-3, noise=0.5)
_, L, _ = S.synthetic_batch(B, H, W, NCLS, seed=5 + rank)
L = L.cuda(); y = L[:, :NCLS].contiguous()
gen = torch.Generator(device='cuda').manual_seed(3 + rank)
hs = (((H + 198) // 2 // 2 // 2) // 2, ((W + 198) // 2 // 2 // 2) // 2)
h = K.pack_nchw(torch.relu(torch.randn((B, 512) + hs, device='cuda', generator=gen)), 512)
nm = torch.randn(y.shape, device='cuda', generator=gen); nk = torch.randn(y.shape, device='cuda', generator=gen)
world = World()

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([s.elapsed_time(e) / n], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

res = {'no exchange': timed(lambda: tr.step(h, y, L, nm, nk)), 'blocking all-reduce': timed(lambda: tr.step(h, y, L, nm, nk, world=world))}
for r in (0, 4, 8, 16, 24):
    tr.DP_RESERVED_SMS = r
    res['bucketed, %d SMs reserved' % r] = timed(lambda: tr.step_dp(h, y, L, nm, nk, world))
if rank == 0:
    for k, v in res.items():
        print('%-32s %.3f ms / step' % (k, v))
dist.destroy_process_group()

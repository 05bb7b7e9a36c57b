"""BASELINE.json configs[3]: DAE train step (rmsprop, crossentropy + squared_error, noise 0.5), batch 10 at 224x224."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S, _kernels as K
from iterative_inference_segm_b200.models.fcn8 import buildFCN8
from iterative_inference_segm_b200.train_dae import DAETrainer
from iterative_inference_segm_b200.profiling import KernelTimer

NCLS, H, W, B = 11, 224, 224, 10
fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=S.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0))
tr = DAETrainer(NCLS, 512, 100, S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1), learning_rate=1e-3, noise=0.5)
X, L, _ = S.synthetic_batch(B, H, W, NCLS, seed=5)
X, L = X.cuda(), L.cuda()
y = L[:, :NCLS].contiguous()
out = fcn[0].net.forward(X, want=('pool4', 'probs_dimshuffle'))
h = out['pool4']
gen = torch.Generator(device='cuda').manual_seed(1)

def step():
    nm = torch.randn(y.shape, device='cuda', generator=gen)
    nk = torch.randn((6,) + tuple(y.shape), device='cuda', generator=gen)      # one mask-noise draw per DePool2D (the reference's graph)
    tr.step(h, y, L, nm, nk)

losses = []
for _ in range(3):
    step(); losses.append(tr.loss_value())
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
n = 5
for _ in range(n):
    step()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
losses.append(tr.loss_value())
def step_g():
    nm = torch.randn(y.shape, device='cuda', generator=gen)
    nk = torch.randn((6,) + tuple(y.shape), device='cuda', generator=gen)      # one mask-noise draw per DePool2D (the reference's graph)
    tr.step_graphed(h, y, L, nm, nk)

for _ in range(3):
    step_g()
torch.cuda.synchronize()
s.record()
for _ in range(n):
    step_g()
e.record(); torch.cuda.synchronize()
ms_g = s.elapsed_time(e) / n
print('graphed train step: %.2f ms / step = %.1f images/s' % (ms_g, B / ms_g * 1e3))
print('train step: %.2f ms / step (batch %d at %dx%d) = %.1f images/s; loss over steps %s; peak mem %.1f GB' % (
    ms, B, H, W, B / ms * 1e3, ['%.4f' % l for l in losses], torch.cuda.max_memory_allocated() / 2**30))
timer = KernelTimer()
import iterative_inference_segm_b200.profiling as P
P._WRAPPED += [n_ for n_ in ('noise_pack', 'loss_grad', 'depool2_bwd', 'pool2_relu_bwd', 'transpose_shift', 'rmsprop_pack') if n_ not in P._WRAPPED]
with timer.recording():
    step()
tot = {}
for (name, tag), v in timer.summary().items():
    tot[name] = tot.get(name, 0.0) + sum(v)
print('kernel sums per step (ms):', {k: round(v, 2) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})

convs = sorted(((sum(v), tag) for (name, tag), v in timer.summary().items() if name == 'conv2d'), reverse=True)[:14]
for ms_, tag in convs:
    print('  conv2d %.3f ms  src %s c1 %s Cout %s RxS %sx%s win %s' % (ms_, tag[0], tag[1], tag[2], tag[3], tag[4], tag[5]))

"""Where one bench step goes: FCN8 forward, the 50-iteration CUDA-graph replay, the eager boundary ops."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S, _kernels as K
from iterative_inference_segm_b200.models.fcn8 import buildFCN8
from iterative_inference_segm_b200.models.DAE_h import buildDAE
from iterative_inference_segm_b200.functions import IterativeInference
from iterative_inference_segm_b200.profiling import KernelTimer

def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

NCLS, H, W, B, N = 11, 360, 480, 10, 50
pf = S.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0); pd = S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf)
dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64,
               additional_pool=2, skip=True, unpool_type='trackind', params=pd)
X, L, lab = S.synthetic_batch(B, H, W, NCLS, seed=100)
X, L = X.cuda(), L.cuda()
fnet = fcn[0].net
ii = IterativeInference(dae, NCLS, [NCLS])
out = fnet.forward(X)
print('fcn8 forward      %.3f ms' % ev(lambda: fnet.forward(X)))
res = ii.run(out['pool4'], out['probs_dimshuffle'], 0.05, N, onehot=L)
print('loop (ii.run)     %.3f ms' % ev(lambda: ii.run(out['pool4'], out['probs_dimshuffle'], 0.05, N, onehot=L)))
st = ii._buffers(B, H, W, N, False)
g = list(st['graph'].values())[0]
t = ev(lambda: g.replay())
print('graph replay only %.3f ms  = %.1f us / iteration' % (t, t / N * 1e3))
timer = KernelTimer()
with timer.recording():
    for _ in range(3):
        ii._loop(st, 0.05, 2, 1e-3, True, False)
tot = {}
for (name, tag), v in timer.summary().items():
    tot[name] = tot.get(name, 0.0) + sum(v[len(v) // 3:]) / (len(v) - len(v) // 3) * (len(v) / 3)
print('eager kernel sums for 2 iterations (1 full + 1 steady):', {k: round(v * 1e3, 1) for k, v in tot.items()}, 'us')
timer = KernelTimer()
with timer.recording():
    for _ in range(3): fnet.forward(X)
tot = {}
for (name, tag), v in timer.summary().items():
    tot[name] = tot.get(name, 0.0) + sum(v) / 3
print('fcn8 kernel sums:', {k: round(v * 1e3, 1) for k, v in tot.items()}, 'us')

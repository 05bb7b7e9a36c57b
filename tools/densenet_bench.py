"""BASELINE.json configs[2]: FC-DenseNet103 + DAE_h iterative inference, batch 10 x 360x480, 50 steps (timing only)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S
from iterative_inference_segm_b200.models.FCDenseNet import build_fcdensenet
from iterative_inference_segm_b200.models.DAE_h import buildDAE
from iterative_inference_segm_b200.functions import IterativeInference
from iterative_inference_segm_b200.profiling import KernelTimer

def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

NCLS, H, W, B, N = 11, 360, 480, 10, 50
fcn = build_fcdensenet(None, ['pool4'], 3, NCLS, params=S.synthetic_densenet_params(3, NCLS, seed=2, logit_gain=4.0))
dae = buildDAE([None], None, NCLS, nb_features_to_concat=464, padding=0, concat_h=['pool4'], noise=0.0, n_filters=64,
               additional_pool=2, skip=True, unpool_type='trackind', params=S.synthetic_dae_params(NCLS, 464, seed=1, out_gain=0.1))
X, L, _ = S.synthetic_batch(B, H, W, NCLS, seed=100)
X, L = X.cuda(), L.cuda()
net = fcn[0].net
ii = IterativeInference(dae, NCLS, [NCLS])
def step():
    out = net.forward(X, want=('pool4', 'probs_dimshuffle'))
    return ii.run(out['pool4_bf16'], out['probs_dimshuffle'], 0.05, N, onehot=L)
t_f = ev(lambda: net.forward(X, want=('pool4', 'probs_dimshuffle')))
t_s = ev(step)
out_ = net.forward(X, want=('pool4', 'probs_dimshuffle'), use_graph=False)
t_l = ev(lambda: ii.run(out_['pool4_bf16'], out_['probs_dimshuffle'], 0.05, N, onehot=L))
t_l0 = ev(lambda: ii.run(out_['pool4_bf16'], out_['probs_dimshuffle'], 0.05, N))
print('loop alone %.2f ms with final metrics, %.2f ms without' % (t_l, t_l0))
res = step(); torch.cuda.synchronize()
print('densenet forward %.2f ms; full step %.2f ms -> %.1f images/s; n_exec %s; peak mem %.1f GB' % (
    t_f, t_s, B / t_s * 1e3, res['n_exec'].cpu().tolist(), torch.cuda.max_memory_allocated() / 2**30))
timer = KernelTimer()
import iterative_inference_segm_b200._kernels as K
with timer.recording():
    net.forward(X, want=('pool4', 'probs_dimshuffle'), use_graph=False)
tot = {}
for (name, tag), v in timer.summary().items():
    tot[name] = tot.get(name, 0.0) + sum(v)
print('forward kernel sums (wrapped launches only):', {k: round(v, 2) for k, v in tot.items()}, 'ms')

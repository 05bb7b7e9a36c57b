"""Profiling target: two eager FC-DenseNet103 forward passes (batch 10, 360x480)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S
from iterative_inference_segm_b200.models.FCDenseNet import build_fcdensenet
fcn = build_fcdensenet(None, ['pool4'], 3, 11, params=S.synthetic_densenet_params(3, 11, seed=2, logit_gain=4.0))
X, _, _ = S.synthetic_batch(10, 360, 480, 11, seed=100)
X = X.cuda()
for _ in range(2):
    fcn[0].net.forward(X, want=('pool4', 'probs_dimshuffle'), use_graph=False)
torch.cuda.synchronize()
print('ok')

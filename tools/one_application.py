"""Profiling target: FCN8 forward + 2 eager loop iterations at the benchmark size (batch 10,
360x480), no CUDA graph, so ncu sees ~100 plain kernel launches in a known order:
  FCN8: pack, 13 conv + 5 pool, fc6, fc7, score_fr, score_pool4, deconv, score_pool3, deconv, deconv, softmax
  per iteration: 6 x (conv, pool), 6 x (unpool, conv), softmax_update, norm_finalize
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights  # noqa: E402  synthetic recipe only


def main(B=10, H=360, W=480, iters=2):
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import IterativeInference
    NCLS = 11
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,
                   n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd)
    X, L, lab = weights.synthetic_batch(B, H, W, NCLS, seed=100)
    out = fcn[0].net.forward(X.cuda(), want=('pool4', 'probs_dimshuffle'))
    ii = IterativeInference(dae, NCLS, [NCLS])
    res = ii.run(out['pool4'], out['probs_dimshuffle'], 0.05, iters, labels=lab.to(torch.int32).cuda(), use_graph=False)
    torch.cuda.synchronize()
    print('ok n_exec', res['n_exec'].cpu().tolist())


if __name__ == '__main__':
    main(*[int(a) for a in sys.argv[1:]])

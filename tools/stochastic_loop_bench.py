"""Times the 50-iteration loop (batch 10, 360x480, mixed) for a DAE built with noise = 0.5: the reference's graph then runs one
noised mask pass per DePool2D in every application (iterative_inference_valid.py's default dae_dict), against noise = 0.
    python tools/stochastic_loop_bench.py"""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from iterative_inference_segm_b200 import synthetic as S  # noqa: E402
from iterative_inference_segm_b200.functions import IterativeInference, function_pred_fcn  # noqa: E402
from iterative_inference_segm_b200.models.DAE_h import buildDAE  # noqa: E402
from iterative_inference_segm_b200.models.fcn8 import buildFCN8  # noqa: E402

NCLS = 11
X, L, lab = S.synthetic_batch(10, 360, 480, NCLS, seed=0)
fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=S.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0),
                precision='mixed')
h, y0 = function_pred_fcn(fcn)(X.cuda())
for noise in (0.0, 0.5):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=noise, n_filters=64,
                       conv_before_pool=1, additional_pool=2, skip=True, unpool_type='trackind',
                       params=S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1), precision='mixed')
    ii = IterativeInference(dae, NCLS, [NCLS])
    labels = lab.to(torch.int32).cuda()
    for _ in range(3):
        ii.run(h, y0, 0.05, 50, eps=0.0, labels=labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ii.run(h, y0, 0.05, 50, eps=0.0, labels=labels)
    e1.record()
    torch.cuda.synchronize()
    print('noise %.1f: %.1f ms per 50-iteration loop of 10 images (%.2f ms per application)' % (noise, e0.elapsed_time(e1) / 3, e0.elapsed_time(e1) / 150))

"""Measures (on a B200) the parity numbers the tests' tolerances are derived from:
FCN8 forward, teacher-forced DAE applications and the free-running loop, CUDA path vs CPU oracle.

    python tools/parity_report.py [H W N_ITER]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets, weights  # noqa: E402  (checker only)


def main(H=64, W=80, N=10, step=0.05, precision='bf16'):
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import (function_pred_fcn, function_pred_dae, IterativeInference)
    NCLS = 11
    X, L, lab = weights.synthetic_batch(2, H, W, NCLS)
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, precision=precision)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,
                   n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision=precision)
    print('precision', precision)
    t = time.time()
    h_o, y0_o = nets.fcn8_forward(pf, X, NCLS)
    print('oracle fcn8 %.1fs' % (time.time() - t))
    h_d, y0_d = function_pred_fcn(fcn)(X.numpy())
    print('FCN8: pool4 max-abs %.3e (scale %.3e)  y0 max-abs %.3e  argmax agree %.5f' % (
        np.abs(h_d - h_o.numpy()).max(), float(h_o.abs().max()), np.abs(y0_d - y0_o.numpy()).max(),
        (y0_d.argmax(1) == y0_o.numpy().argmax(1)).mean()))
    # teacher-forced and free-running
    pred_dae = function_pred_dae(dae)
    y_o = y0_o.clone()
    ii = IterativeInference(dae, NCLS, [NCLS])
    res = ii.run(torch.from_numpy(h_o.numpy()).cuda(), y0_o.cuda(), step, N, eps=0.0,
                 labels=lab.to(torch.int32).cuda(), per_iter_metrics=True, use_graph=False)
    y_dev_iters = None
    y_free = y0_o.cuda().clone()
    ii1 = IterativeInference(dae, NCLS, [NCLS])
    for it in range(N):
        p_o = nets.dae_forward(pd, y_o, h_o, 100)
        p_tf = pred_dae(h_o.numpy(), y_o.numpy())                      # teacher-forced: oracle y in
        y_o = torch.clamp(y_o - step * (y_o - p_o), 0, 1)
        y_free = ii1.run(torch.from_numpy(h_o.numpy()).cuda(), y_free, step, 1, eps=0.0, use_graph=False)['y'].clone()
        print('it %2d  teacher-forced p max-abs %.3e mean-abs %.3e | free-running y max-abs %.3e argmax agree %.5f' % (
            it + 1, np.abs(p_tf - p_o.numpy()).max(), np.abs(p_tf - p_o.numpy()).mean(),
            float((y_free.cpu() - y_o).abs().max()), float((y_free.cpu().argmax(1) == y_o.argmax(1)).float().mean())),
            flush=True)
    print('loop (one call, N=%d): y max-abs %.3e  n_exec %s' % (
        N, float((res['y'].cpu() - y_o).abs().max()), res['n_exec'].cpu().tolist()))


if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:4]]
    main(*a, precision=sys.argv[4] if len(sys.argv) > 4 else 'bf16')

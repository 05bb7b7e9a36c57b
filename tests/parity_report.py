"""Measures (on a B200) the parity numbers the tests' tolerances are derived from, CUDA path vs the CPU oracle
(test infrastructure; not collected by pytest):

  * FCN8 forward (pool4, y0);
  * the TRUE pipeline: device FCN8 -> device loop (one CUDA-graph replay, y recorded after every iteration)
    against oracle FCN8 -> oracle loop, per iteration;
  * teacher-forced DAE applications: the oracle's (h, y_k) in, p_k compared;
  * the loop alone from the oracle's h / y0 (isolates the DAE path from FCN8's error).

    python tests/parity_report.py H W N_ITER [precision] [n_images]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets, weights  # noqa: E402  (checker only)


def measure(H=64, W=80, N=10, step=0.05, precision='bf16', n_images=2, verbose=True):
    """Returns a dict of worst-case numbers over the N iterations."""
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import (function_pred_fcn, function_pred_dae, IterativeInference)
    NCLS = 11
    say = print if verbose else (lambda *a, **k: None)
    X, L, lab = weights.synthetic_batch(n_images, H, W, NCLS)
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, precision=precision)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,
                   n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision=precision)
    say('precision', precision, '%dx%d' % (H, W), 'N', N, 'images', n_images)
    t = time.time()
    h_o, y0_o = nets.fcn8_forward(pf, X, NCLS)
    say('oracle fcn8 %.1fs' % (time.time() - t))
    h_d, y0_d = function_pred_fcn(fcn)(X.cuda())
    out = {'fcn_y0_maxabs': float((y0_d.cpu() - y0_o).abs().max()),
           'fcn_argmax': float((y0_d.cpu().argmax(1) == y0_o.argmax(1)).float().mean()),
           'fcn_pool4_rel': float((h_d.cpu() - h_o).abs().max() / h_o.abs().max())}
    say('FCN8: pool4 max-abs / scale %.3e  y0 max-abs %.3e  argmax agree %.5f' % (out['fcn_pool4_rel'], out['fcn_y0_maxabs'], out['fcn_argmax']))
    ii = IterativeInference(dae, NCLS, [NCLS])
    # the true pipeline: device FCN8 -> device loop, one graph replay
    true_y = ii.run(h_d, y0_d, step, N, eps=0.0, record_y=True)['y_hist'].cpu().clone()
    # the loop alone: oracle h / y0 in
    loop_y = ii.run(h_o.cuda(), y0_o.cuda(), step, N, eps=0.0, record_y=True)['y_hist'].cpu().clone()
    pred_dae = function_pred_dae(dae)
    y_o = y0_o.clone()
    worst = {'tf_p': 0.0, 'true_y': 0.0, 'true_argmax': 1.0, 'loop_y': 0.0, 'loop_argmax': 1.0}
    for it in range(N):
        p_o = nets.dae_forward(pd, y_o, h_o, 100)
        p_tf = pred_dae(h_o.cuda(), y_o.cuda()).cpu()                   # teacher-forced: oracle y in
        y_o = torch.clamp(y_o - step * (y_o - p_o), 0, 1)
        am_o = y_o.argmax(1)
        r = {'tf_p': float((p_tf - p_o).abs().max()), 'true_y': float((true_y[it] - y_o).abs().max()),
             'true_argmax': float((true_y[it].argmax(1) == am_o).float().mean()),
             'loop_y': float((loop_y[it] - y_o).abs().max()), 'loop_argmax': float((loop_y[it].argmax(1) == am_o).float().mean())}
        for k, v in r.items():
            worst[k] = min(worst[k], v) if 'argmax' in k else max(worst[k], v)
        say('it %2d  teacher-forced p %.3e | true pipeline y %.3e argmax %.5f | loop alone y %.3e argmax %.5f' % (
            it + 1, r['tf_p'], r['true_y'], r['true_argmax'], r['loop_y'], r['loop_argmax']), flush=True)
    out.update(worst)
    say('worst over %d iterations: %s' % (N, {k: ('%.3e' % v if 'argmax' not in k else '%.5f' % v) for k, v in worst.items()}))
    return out


if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:4]]
    measure(*a, precision=sys.argv[4] if len(sys.argv) > 4 else 'bf16', n_images=int(sys.argv[5]) if len(sys.argv) > 5 else 2)

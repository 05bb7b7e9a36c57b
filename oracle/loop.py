"""The iterative-inference host loop restated (test oracle).

Follows iterative_inference.py:258-291 and iterative_inference_valid.py:265-297.
"""
import numpy as np
import torch

from . import metrics as M
from .nets import dae_forward

EPSILON = 1e-3  # iterative_inference.py:53 (the second _EPSILON wins)


def de_fn(params, h, y, padding, **dae_kw):
    """de = -(pred_dae - y)  (iterative_inference.py:203-204)."""
    return y - dae_forward(params, y, h, padding, **dae_kw)


def iterate_image(params, h_im, y_im, step, num_iter, padding, eps=EPSILON,
                  t_im=None, n_classes=None, void_labels=(), record=False,
                  forward=None, **dae_kw):
    """Per-image loop (iterative_inference.py:265-280):
        grad = de_fn(h, y); y = clip(y - step*grad, 0, 1)
        norm = mean_{b,h,w} ||grad||_2 over channels; break if norm < eps
        (the break is after the update and before that iteration's val_fn).
    Returns (y, n_executed, per_iter) where per_iter[it] = (acc, jacc, mse) of
    the iterations that reached val_fn; `record` additionally keeps y and p.
    `forward(y, h) -> p` replaces DAE_h by another DAE kind (contextmod, fcn8: iterative_inference.py:165-179)."""
    y = y_im.clone()
    per_iter, trace = [], []
    n_exec = 0
    for it in range(num_iter):
        p = forward(y, h_im) if forward is not None else dae_forward(params, y, h_im, padding, **dae_kw)
        grad = y - p
        y = torch.clamp(y - step * grad, 0.0, 1.0)
        n_exec += 1
        norm = float(torch.linalg.vector_norm(grad, dim=1).mean())
        if record:
            trace.append({'p': p.clone(), 'y': y.clone(), 'norm': norm})
        if norm < eps:
            break
        if t_im is not None:
            per_iter.append(M.val_fn(y.numpy(), t_im, n_classes, void_labels))
    return y, n_exec, per_iter, trace


def inference_batch(params_dae, H, Y, step, num_iter, padding, L=None,
                    n_classes=None, void_labels=(), eps=EPSILON, forward=None, **dae_kw):
    """One batch of iterative_inference.py:258-291: per-image loops, then the
    batch-level val_fn on the concatenated result.  Also accumulates the
    valid-script matrix valid_mat[:, :, it] += jacc_iter
    (iterative_inference_valid.py:231,288)."""
    B = Y.shape[0]
    outs, n_execs = [], []
    valid_mat = np.zeros((2, n_classes, num_iter)) if L is not None else None
    for im in range(B):
        t_im = None if L is None else L[im:im + 1]
        y, n_exec, per_iter, _ = iterate_image(
            params_dae, H[im:im + 1], Y[im:im + 1], step, num_iter, padding,
            eps=eps, t_im=t_im, n_classes=n_classes, void_labels=void_labels,
            forward=forward, **dae_kw)
        outs.append(y)
        n_execs.append(n_exec)
        if L is not None:
            for it, (_, jacc_iter, _) in enumerate(per_iter):
                valid_mat[:, :, it] += jacc_iter
    Y_ii = torch.cat(outs, dim=0)
    batch_metrics = None
    if L is not None:
        batch_metrics = M.val_fn(Y_ii.numpy(), L, n_classes, void_labels)
    return Y_ii, n_execs, batch_metrics, valid_mat

"""metrics.py of the reference on the sm_100a metrics kernel (one pass, int64 counts).

Same names and argument meaning as the reference (`metrics.py:11-65,144-156`); inputs may be
the 2-D `(N, C)` arrays the reference builds with dimshuffle+reshape
(`iterative_inference.py:193-200`) or 4-D NCHW tensors, numpy or torch.  Outputs are float32
like the reference's (`cm = T.zeros(...)` is floatX, `metrics.py:23`); `confusion_matrix` exposes
the exact int64 counts the float views are derived from.
"""
import numpy as np
import torch

from . import _kernels as K
from .functions import MetricsAccumulator, jaccard_from_cm, _to_cuda


def _as_nchw(a):
    t, _ = _to_cuda(a)
    if t.dim() == 2:            # (N, C) rows = pixels  ->  (1, C, N, 1)
        t = t.t().contiguous().view(1, t.shape[1], t.shape[0], 1)
    return t


def _accumulate(y_pred, y_true, n_classes, void_label):
    y = _as_nchw(y_pred)
    t = _as_nchw(y_true)
    acc = MetricsAccumulator(y.shape[0], n_classes)
    if t.shape[1] == n_classes + 1:
        K.metrics_accumulate(y, acc.cm, acc.counts, acc.sqerr, onehot=t, void_label=void_label)
    else:
        raise ValueError('target must be one-hot with n_classes+1 channels (last = void)')
    return acc


def confusion_matrix(y_pred, y_true, n_classes):
    """int64 (C, C): cm[pred, true] over pixels whose true class is < n_classes (metrics.py:22-27)."""
    acc = _accumulate(y_pred, y_true, n_classes, n_classes)
    return acc.cm.sum(0).cpu().numpy().reshape(n_classes, n_classes)


def jaccard(y_pred, y_true, n_classes, one_hot=False):
    """metrics.py:11-37 -> float32 (2, C) = [TP; TP+FP+FN]."""
    if not one_hot:
        raise NotImplementedError('the iterative-inference path always passes one_hot=True')
    return jaccard_from_cm(confusion_matrix(y_pred, y_true, n_classes))


def accuracy(y_pred, y_true, void_labels, one_hot=False):
    """metrics.py:40-65 -> float32 scalar."""
    if not one_hot:
        raise NotImplementedError('the iterative-inference path always passes one_hot=True')
    n_classes = _as_nchw(y_pred).shape[1]
    vl = int(void_labels[0]) if len(void_labels) else -1
    acc = _accumulate(y_pred, y_true, n_classes, vl)
    c = acc.counts.sum(0).cpu().numpy()
    return np.float32(np.float32(c[0]) / np.float32(c[1]))


def squared_error(y_pred, y_true, void):
    """metrics.py:144-156 with integer `void` (= n_classes, iterative_inference.py:125)."""
    if not isinstance(void, int):
        raise NotImplementedError('list-valued void is not used on the iterative-inference path')
    acc = _accumulate(y_pred, y_true, void, void)
    s = acc.sqerr.sum(0).cpu().numpy()
    return np.float32(s[0] / s[1])

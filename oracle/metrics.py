"""metrics.py of the reference restated in numpy (test oracle).

Inputs follow val_fn (iterative_inference.py:193-210): y (B,C,H,W) float32
probabilities, target (B,C+1,H,W) one-hot float32 whose last channel is void.
"""
import numpy as np


def _to_2d(a):
    """dimshuffle((0,2,3,1)).reshape(N, C)  (iterative_inference.py:193-200)."""
    a = np.asarray(a)
    return a.transpose(0, 2, 3, 1).reshape(-1, a.shape[1])


def confusion_matrix(y, target, n_classes):
    """int64 cm[i, j] = #{pred == i and true == j}, i, j < n_classes
    (metrics.py:22-27; rows = prediction, cols = truth).  argmax ties -> first
    index (T.argmax)."""
    pred = np.argmax(_to_2d(y), axis=1)
    true = np.argmax(_to_2d(target), axis=1)
    keep = true < n_classes
    idx = pred[keep] * n_classes + true[keep]
    return np.bincount(idx, minlength=n_classes * n_classes).reshape(
        n_classes, n_classes).astype(np.int64)


def jaccard_from_cm(cm):
    """metrics.py:29-37: stack([TP, TP+FP+FN]) as float32; FP = row sums - TP,
    FN = column sums - TP."""
    tp = np.diag(cm).astype(np.float32)
    fp = cm.sum(1).astype(np.float32) - tp
    fn = cm.sum(0).astype(np.float32) - tp
    return np.stack([tp, tp + fp + fn], axis=0).astype(np.float32)


def jaccard(y, target, n_classes):
    """metrics.py:11-37 with one_hot=True."""
    return jaccard_from_cm(confusion_matrix(y, target, n_classes))


def accuracy_counts(y, target, void_labels):
    pred = np.argmax(_to_2d(y), axis=1)
    true = np.argmax(_to_2d(target), axis=1)
    mask = np.ones_like(true, dtype=bool)
    for el in void_labels:
        mask &= true != el
    return int(((pred == true) & mask).sum()), int(mask.sum())


def accuracy(y, target, void_labels):
    """metrics.py:40-65: sum(eq * mask) / sum(mask), float32."""
    c, v = accuracy_counts(y, target, void_labels)
    return np.float32(np.float32(c) / np.float32(v))


def squared_error(y, target, void):
    """metrics.py:144-156 with integer `void` (= n_classes,
    iterative_inference.py:125): mean over channels of (y - t[:, :void])^2,
    masked by sum_c t[:, :void], divided by the mask sum."""
    y = np.asarray(y, dtype=np.float32)
    t = np.asarray(target, dtype=np.float32)[:, :void]
    loss_aux = ((y - t) ** 2).mean(axis=1)
    mask = t.sum(axis=1)
    return np.float32((loss_aux * mask).sum(dtype=np.float32) / mask.sum(dtype=np.float32))


def crossentropy(y_2d, target_2d, void_labels):
    """metrics.py:68-91 with one_hot=True (train_dae.py:279-286): clip to
    [1e-7, 1-1e-7], categorical CE against argmax(target) with void pixels
    re-labelled 0 and masked out of the mean."""
    eps = 10e-8
    p = np.clip(np.asarray(y_2d, dtype=np.float32), eps, 1.0 - eps)
    true = np.argmax(np.asarray(target_2d), axis=1)
    mask = np.ones_like(true, dtype=np.float32)
    for el in void_labels:
        mask[true == el] = 0.
    tt = (true * mask).astype(np.int64)
    loss = -np.log(p[np.arange(p.shape[0]), tt])
    return np.float32((loss * mask).sum(dtype=np.float32) / mask.sum(dtype=np.float32))


def val_fn(y, target, n_classes, void_labels):
    """[acc, jacc(2,C), mse] as in iterative_inference.py:207-210."""
    return (accuracy(y, target, void_labels), jaccard(y, target, n_classes),
            squared_error(y, target, n_classes))


def print_results_values(rec, acc, jacc, nbatches):
    """helpers.py:172-177: (loss/nbatches, acc/nbatches, nanmean(num/denom))."""
    with np.errstate(divide='ignore', invalid='ignore'):
        jm = np.nanmean(jacc[0, :] / jacc[1, :])
    return rec / nbatches, acc / nbatches, jm

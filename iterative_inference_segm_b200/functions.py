"""The compiled callables of the reference, on the sm_100a kernels.

The reference compiles four Theano functions (iterative_inference.py:187-210):
    pred_fcn_fn(X)        -> [h..., y0]
    pred_dae_fn(*h, y)    -> DAE(y, h)
    de_fn(*h, y)          -> y - DAE(y, h)
    val_fn(y, target)     -> [acc, jacc(2,C) float32, mse]
and drives them from a Python loop that crosses the device boundary 2-3 times per
iteration per image.  This module provides the same four callables (numpy or
torch in, same kind out) and `IterativeInference`, which moves the whole loop
(iterative_inference.py:258-291) onto the device: h and y stay resident in HBM,
N iterations of {DAE, softmax+update, norm test, optional per-iteration metrics}
are captured once as a CUDA Graph and replayed.
"""
import numpy as np
import torch

from . import _kernels as K

EPSILON = 1e-3   # iterative_inference.py:53


def _to_cuda(a, dtype=torch.float32):
    """numpy / torch (any device) -> contiguous CUDA tensor; remembers whether to return numpy."""
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a)).to('cuda', dtype=dtype, non_blocking=True), True
    return a.to('cuda', dtype=dtype, non_blocking=True).contiguous(), False


def _ret(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def function_pred_fcn(fcn):
    """theano.function([input_x_var], get_output(fcn, deterministic=True))
    (iterative_inference.py:187-188).  `fcn` is the handle list of buildFCN8."""
    net = fcn[0].net
    names = [hd.name for hd in fcn]

    def pred_fcn_fn(X):
        Xc, as_np = _to_cuda(X)
        out = net.forward(Xc, want=tuple(names))
        res = []
        for hd in fcn:
            t = out[hd.name]
            if hd.name == 'input':
                t = t.clone()
            elif hd.name.startswith('pool'):    # poolK features (NHWC bf16 / fp32 inside) leave as the reference's NCHW float32
                t = K.unpack_nhwc(t, hd.output_shape[1], split=net.split and t.dtype == torch.bfloat16)
            res.append(_ret(t, as_np))
        return res
    return pred_fcn_fn


def _pack_h(net, h, out=None):
    """The conditioning tensor in the layout the DAE reads: NHWC bf16 (pairs) for the tensor-core DAE_h; nets with their
    own layout (the fp32 context module) provide pack_h."""
    if hasattr(net, 'pack_h'):
        return net.pack_h(h, out)
    return K.pack_nchw(h.contiguous(), net.h_pad, out=out, split=net.split)


class _DaeCallable(object):
    def __init__(self, dae, grad):
        self.net = dae.net
        self.grad = grad

    def __call__(self, *args):
        h, y = args[0], args[-1]
        assert len(args) == 2, 'one conditioning tensor (concat_h has one entry) + y'
        net = self.net
        hc, _ = _to_cuda(h)
        yc, as_np = _to_cuda(y)
        h_bf16 = _pack_h(net, hc)
        y_bf16 = K.pack_nchw(yc, net.y_cpad, split=net.split)
        logits = net.logits(h_bf16, y_bf16, y_f32=yc)
        out = torch.empty_like(yc)
        if self.grad:
            K.softmax_grad(logits, yc, out)
        else:
            K.softmax_nchw(logits, net.n_classes, out)
        return _ret(out, as_np)


def function_pred_dae(dae):
    """theano.function(h_vars + [y_hat_var], get_output(dae, deterministic=True))
    (iterative_inference.py:189-190)."""
    return _DaeCallable(dae, grad=False)


def function_de(dae):
    """de = -(pred_dae - y_hat_var)  (iterative_inference.py:203-204)."""
    return _DaeCallable(dae, grad=True)


class MetricsAccumulator(object):
    """Device-side int64 confusion matrix + accuracy counts + squared-error sums."""

    def __init__(self, n_images, n_classes, device='cuda'):
        self.C = n_classes
        self.cm = torch.zeros((n_images, n_classes * n_classes), dtype=torch.int64, device=device)
        self.counts = torch.zeros((n_images, 2), dtype=torch.int64, device=device)
        self.sqerr = torch.zeros((n_images, 2), dtype=torch.float64, device=device)

    def zero_(self):
        self.cm.zero_(); self.counts.zero_(); self.sqerr.zero_()


def jaccard_from_cm(cm):
    """metrics.py:29-37 on an integer confusion matrix (rows = prediction): float32 (2, C)."""
    cm = np.asarray(cm).reshape(int(round(np.sqrt(cm.size))), -1)
    tp = np.diag(cm).astype(np.float32)
    fp = cm.sum(1).astype(np.float32) - tp
    fn = cm.sum(0).astype(np.float32) - tp
    return np.stack([tp, tp + fp + fn], axis=0).astype(np.float32)


def _void_label(n_classes, void_labels):
    if len(void_labels) == 0:
        return -1
    assert len(void_labels) == 1, 'one void label supported'
    return int(void_labels[0])


def function_val(n_classes, void_labels):
    """theano.function([y_hat_var, target_var], [test_acc, test_jacc, test_loss])
    (iterative_inference.py:207-210): accuracy (metrics.py:40-65), jaccard (:11-37),
    squared_error with void = n_classes (:144-156), all float32 like the reference."""
    vl = _void_label(n_classes, void_labels)

    def val_fn(y, target):
        yc, _ = _to_cuda(y)
        tc, _ = _to_cuda(target)
        acc = MetricsAccumulator(yc.shape[0], n_classes)
        K.metrics_accumulate(yc, acc.cm, acc.counts, acc.sqerr, onehot=tc, void_label=vl)
        cm = acc.cm.sum(0).cpu().numpy()
        cnt = acc.counts.sum(0).cpu().numpy()
        se = acc.sqerr.sum(0).cpu().numpy()
        with np.errstate(divide='ignore', invalid='ignore'):
            a = np.float32(np.float32(cnt[0]) / np.float32(cnt[1]))
            mse = np.float32(se[0] / se[1])
        return [a, jaccard_from_cm(cm), mse]
    return val_fn


class IterativeInference(object):
    """The loop of iterative_inference.py:258-291 for a whole batch, on the device.

    Per iteration k, for every image still active:
        p = DAE(y, h);  g = y - p;  y <- clip(y - step*g, 0, 1)
        norm = mean_pixels ||g||_2;  n_exec += 1;  if norm < eps: active = 0
    (update first, then the test, then -- only if still active -- that iteration's
    metrics: iterative_inference.py:267-280).  Frozen images are predicated off."""

    def __init__(self, dae, n_classes=None, void_labels=(), fuse_update=True):
        self.net = dae.net
        self.fuse_update = fuse_update      # False: stand-alone softmax_update kernel (same bits, one more pass over HBM)
        self.C = self.net.n_classes if n_classes is None else n_classes
        self.void_label = _void_label(self.C, list(void_labels))
        self._state = {}
        self.graph_kernel_nodes = 0
        self.graph_captures = 0

    def _buffers(self, B, H, W, num_iter, per_iter, record_y=False):
        key = (B, H, W, num_iter, per_iter, record_y)
        st = self._state.get(key)
        if st is None:
            dev = self.net.device
            hs = self.net.h_spatial(H, W)
            st = {
                'h': self.net.alloc_h(B, H, W) if hasattr(self.net, 'alloc_h') else
                     torch.zeros((B,) + hs + (self.net.cm * self.net.h_pad,), dtype=torch.bfloat16, device=dev),
                'y': torch.zeros((B, self.C, H, W), dtype=torch.float32, device=dev),
                'y_bf16': torch.zeros((B, H, W, self.net.cm * self.net.y_cpad), dtype=torch.bfloat16, device=dev),
                'labels': torch.zeros((B, H, W), dtype=torch.int32, device=dev),
                'active': torch.ones((B,), dtype=torch.int32, device=dev),
                'n_exec': torch.zeros((B,), dtype=torch.int32, device=dev),
                'norm': torch.zeros((B,), dtype=torch.float32, device=dev),
                'norm_hist': torch.zeros((num_iter, B), dtype=torch.float32, device=dev),
                'partial': torch.zeros((B, K.update_blocks(H, W)), dtype=torch.float32, device=dev),
                'norm_acc': torch.zeros((B,), dtype=torch.int64, device=dev),
                'step': torch.zeros((1,), dtype=torch.float32, device=dev),      # device scalars: one captured graph
                'eps': torch.zeros((1,), dtype=torch.float32, device=dev),       # serves every (step, eps)
                'final': MetricsAccumulator(B, self.C, dev),
                'iter': [MetricsAccumulator(B, self.C, dev) for _ in range(num_iter)] if per_iter else None,
                'y_hist': torch.zeros((num_iter, B, self.C, H, W), dtype=torch.float32, device=dev) if record_y else None,
                'graph': {},
            }
            self._state[key] = st
        return st

    def _loop(self, st, step, num_iter, eps, with_metrics, per_iter):
        net = self.net
        H, W = st['y'].shape[2:]
        st['active'].fill_(1)
        st['n_exec'].zero_()
        st['norm_acc'].zero_()
        st['final'].zero_()
        if per_iter:
            for acc in st['iter']:
                acc.zero_()
        # bn=1 + DePool2D: the reference iterates one image at a time, so its mask pass sees per-image batch statistics
        stats_kw = {'per_image_stats': True} if getattr(net, 'bn_batch_masks', False) else {}
        for it in range(num_iter):
            # the first iteration of a batch computes the whole contracting path (h is new); later ones only
            # its y-dependent windows -- everything outside them is iteration-invariant (DAENet.down_windows)
            if self.fuse_update and net.fusable_update:
                # softmax tail + update + norm in the epilogue of up_conv1: the logits never reach HBM
                net.logits(st['h'], st['y_bf16'], full_down=(it == 0), y_f32=st['y'],
                           update=dict(y=st['y'], active=st['active'], norm_acc=st['norm_acc'], step=step), **stats_kw)
                K.norm_finalize_fixed(st['norm_acc'], st['norm'], st['active'], st['n_exec'], H, W, eps)
            else:
                kw = {'active': st['active']} if getattr(net, 'takes_active', False) else {}
                kw.update(stats_kw)
                logits = net.logits(st['h'], st['y_bf16'], full_down=(it == 0), y_f32=st['y'], **kw)
                K.softmax_update(logits, st['y'], st['y_bf16'], st['active'], st['partial'], step, split=net.split)
                K.norm_finalize(st['partial'], st['norm'], st['active'], st['n_exec'], H, W, eps)
            st['norm_hist'][it].copy_(st['norm'])
            if st['y_hist'] is not None:
                st['y_hist'][it].copy_(st['y'])
            if per_iter:
                acc = st['iter'][it]
                K.metrics_accumulate(st['y'], acc.cm, acc.counts, acc.sqerr, labels=st['labels'],
                                     active=st['active'], void_label=self.void_label)
        if with_metrics:
            acc = st['final']
            K.metrics_accumulate(st['y'], acc.cm, acc.counts, acc.sqerr, labels=st['labels'],
                                 void_label=self.void_label)

    def run(self, h, y0, step, num_iter, eps=EPSILON, labels=None, onehot=None, per_iter_metrics=False,
            use_graph=True, record_y=False):
        """h: NHWC bf16 (internal, from FCN8Net.forward) or NCHW fp32; y0: NCHW fp32;
        labels: int (B,H,W) class indices with void = the void label, or onehot: the reference's
        (B,C+1,H,W) float32 target (argmax'd once on the device), or neither.
        Returns a dict of device tensors: y (B,C,H,W), n_exec, norm_hist, cm / counts /
        sqerr of the final batch-level val_fn, and the per-iteration accumulators.  `record_y` (parity tooling):
        also keeps y after every iteration in 'y_hist' (num_iter, B, C, H, W)."""
        B, Cc, H, W = y0.shape
        assert Cc == self.C
        with_metrics = labels is not None or onehot is not None
        per_iter = bool(per_iter_metrics and with_metrics)
        st = self._buffers(B, H, W, num_iter, per_iter, bool(record_y))
        if h.dtype == torch.bfloat16:
            st['h'].copy_(h)
        else:
            _pack_h(self.net, h, st['h'])
        st['y'].copy_(y0)
        K.pack_nchw(st['y'], self.net.y_cpad, out=st['y_bf16'], split=self.net.split)
        if onehot is not None:
            K.onehot_to_labels(onehot.contiguous(), st['labels'])
        elif labels is not None:
            st['labels'].copy_(labels)
        if use_graph:
            # step and eps are device scalars read by the kernels at run time: the graph is captured once per shape and
            # replayed for any step size / threshold (the seven step values of the valid sweep share it)
            st['step'].fill_(float(step))
            st['eps'].fill_(float(eps))
            step, eps = st['step'], st['eps']
            gkey = (with_metrics,)
            g = st['graph'].get(gkey)
            if g is None:
                # warm-up outside capture (lazy module loading, smem attribute calls) with every launch shape of the
                # captured loop: the full first iteration, a windowed one and the per-iteration metrics; then capture
                self._loop(st, step, min(num_iter, 2), eps, with_metrics, per_iter)
                torch.cuda.synchronize()
                st['y'].copy_(y0)
                K.pack_nchw(st['y'], self.net.y_cpad, out=st['y_bf16'], split=self.net.split)
                g = torch.cuda.CUDAGraph()
                from . import _lib
                n0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    self._loop(st, step, num_iter, eps, with_metrics, per_iter)
                self.graph_kernel_nodes = _lib.launch_count() - n0     # library kernels per replay
                self.graph_captures += 1
                st['graph'][gkey] = g
            g.replay()
        else:
            self._loop(st, step, num_iter, eps, with_metrics, per_iter)
        return {'y': st['y'], 'n_exec': st['n_exec'], 'norm_hist': st['norm_hist'],
                'cm': st['final'].cm, 'counts': st['final'].counts, 'sqerr': st['final'].sqerr,
                'iter': st['iter'], 'y_hist': st['y_hist']}

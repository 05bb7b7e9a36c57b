"""CUDA-event timing of individual kernel launches (measurement tooling for bench.py).

`KernelTimer` temporarily wraps the launch functions of `_kernels` so that every launch made
through them is bracketed by a pair of CUDA events recorded on the launching stream (torch's
current stream -- the same stream the C ABI receives).  Only for eager (non-graph) passes.
"""
import contextlib

import torch

from . import _kernels as K

_WRAPPED = ['conv2d', 'wgrad_gemm', 'gemm_nt_splitk', 'bias_grad', 'maxpool2', 'unpool2', 'softmax_update', 'softmax_nchw', 'norm_finalize',
            'metrics_accumulate', 'deconv16', 'pack_nchw', 'norm_finalize_fixed', 'bn_relu_pack', 'channel_stats',
            'maxpool2_f32', 'deconv_interleave']


class KernelTimer(object):
    def __init__(self):
        self.records = []   # (name, tag, start_event, end_event)

    @contextlib.contextmanager
    def recording(self):
        saved = {n: getattr(K, n) for n in _WRAPPED}

        def wrap(name, fn):
            def timed(*args, **kw):
                s = torch.cuda.Event(enable_timing=True)
                e = torch.cuda.Event(enable_timing=True)
                s.record()
                out = fn(*args, **kw)
                e.record()
                tag = _tag(name, args, kw)
                if name == 'conv2d':
                    tag = tag + (K.last_conv_plan(),)       # (kernel, BN, KB): which kernel ran
                self.records.append((name, tag, s, e))
                return out
            return timed
        for n, fn in saved.items():
            setattr(K, n, wrap(n, fn))
        try:
            yield self
        finally:
            for n, fn in saved.items():
                setattr(K, n, fn)

    def summary(self):
        """{(name, tag): [ms, ...]} after a device synchronize."""
        torch.cuda.synchronize()
        out = {}
        for name, tag, s, e in self.records:
            out.setdefault((name, tag), []).append(s.elapsed_time(e))
        return out


def _tag(name, args, kw):
    if name == 'conv2d':
        src0, weight = args[0], args[1]
        R, S = args[3], args[4]
        out = kw.get('out')
        win = kw.get('window')
        c1 = kw['src1'].shape[3] if kw.get('src1') is not None else 0
        return (tuple(src0.shape), c1, weight.shape[0], R, S, tuple(win) if win else None,
                tuple(out.shape) if out is not None else None, kw.get('pooled') is not None)
    if name == 'unpool2':
        out = kw.get('out')
        return (tuple(args[0].shape), tuple(out.shape) if out is not None else None)
    if name == 'maxpool2':
        return tuple(args[0].shape)
    return None

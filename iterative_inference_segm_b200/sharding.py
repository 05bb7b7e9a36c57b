"""Image sharding across GPUs and the one collective of the inference path.

The reference is single-process.  Here images are partitioned into contiguous shards (rank r
owns [r*n/G, (r+1)*n/G)), weights are replicated, and the only exchange is an all-reduce (SUM)
of the integer confusion matrix / counts (and the fp64 squared-error sums).  Integer sums are
order-independent, so any number of ranks gives bit-identical matrices.  Works on NCCL (CUDA
tensors) and gloo (CPU tensors)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` for `rank` of `world` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_metrics(cm, counts, sqerr=None):
    """In-place SUM all-reduce of int64 `cm`, int64 `counts` and (optionally) fp64 `sqerr`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return cm, counts, sqerr
    packed = torch.cat([cm.reshape(-1), counts.reshape(-1)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    cm.copy_(packed[:cm.numel()].view_as(cm))
    counts.copy_(packed[cm.numel():].view_as(counts))
    if sqerr is not None:
        dist.all_reduce(sqerr, op=dist.ReduceOp.SUM)
    return cm, counts, sqerr


class World(object):
    """The collectives of the data-parallel DAE training step (train_dae.py step, config 4): SUM all-reduce of the
    loss denominators and of the per-layer weight-gradient matrices.  NCCL for CUDA tensors, gloo for CPU."""

    def __init__(self):
        self.size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.size > 1 else 0

    def allreduce_sum(self, t):
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def allreduce_sum_async(self, t):
        """Starts the SUM all-reduce of `t` (ordered after the kernels already queued on the current stream) and returns
        a handle with `.wait()`; the reduction runs on the backend's own stream, under whatever is launched next."""
        if self.size > 1:
            return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
        return _Done()


class _Done(object):
    def wait(self):
        return True

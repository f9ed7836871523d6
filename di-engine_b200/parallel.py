"""Data-parallel plumbing for the learner hot path.

Every operator on this path is independent per batch column (recurrences run along T only; the n-step / C51 heads are
per sample), so N GPUs simply take N contiguous column shards; per-sample gradients never leave their GPU.  The only
cross-column coupling is that the losses are means over all elements: one all-reduce of a packed vector of the
per-rank loss scalars, divided by the world size -- the reference's DDP semantics of "mean of equal-sized rank means"
(ding/utils/pytorch_ddp_dist_helper.py:38-47: all_reduce then div_(world_size)).

Works with any torch.distributed backend: NCCL over NVLink on the GPUs, gloo on CPU tensors in the tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous share of n columns; requires an even split (as DDP batch sharding does)."""
    if n % world != 0:
        raise ValueError("cannot shard %d columns evenly over %d ranks" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def shard_columns(x, rank, world, dim):
    """Contiguous shard of ``x`` along ``dim`` (the batch axis B of a (T, B, ...) or (B, ...) tensor)."""
    if x is None:
        return None
    lo, hi = shard_bounds(x.shape[dim], rank, world)
    return x.narrow(dim, lo, hi - lo).contiguous()


def shard_trajectory_batch(batch, rank, world, time_major_keys, sample_major_keys):
    """Shard a dict of tensors: ``time_major_keys`` are (T, B, ...) (split on dim 1), ``sample_major_keys`` (B, ...)."""
    out = {}
    for k, v in batch.items():
        if k in time_major_keys:
            out[k] = shard_columns(v, rank, world, 1)
        elif k in sample_major_keys:
            out[k] = shard_columns(v, rank, world, 0)
        else:
            out[k] = v
    return out


class LossAllReduce:
    """One collective per step for all loss scalars of that step.

    ``reduce(values)`` packs 0-dim tensors (or floats) into a preallocated vector, all-reduces it (sum) and divides by
    the world size.  With ``async_op=True`` the handle is returned so the caller can overlap it with the next step.
    """

    def __init__(self, n_values, device, group=None):
        self.buf = torch.zeros(n_values, dtype=torch.float32, device=device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def pack(self, values):
        for i, v in enumerate(values):
            if isinstance(v, torch.Tensor):
                self.buf[i].copy_(v.detach().reshape(()))
            else:
                self.buf[i] = float(v)
        return self.buf

    def reduce(self, values=None, async_op=False):
        if values is not None:
            self.pack(values)
        if self.world == 1:
            return self.buf if not async_op else None
        work = dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        self.buf.div_(self.world)
        return self.buf

    def finish(self, work):
        if work is not None:
            work.wait()
        if self.world > 1:
            self.buf.div_(self.world)
        return self.buf

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


class P2PLossAllReduce:
    """Same contract as ``LossAllReduce`` (mean of the per-rank loss scalars, one exchange per step) but through ONE small
    kernel over NVLink peer memory (``b200rl_p2p_allreduce_mean``) instead of a small-message NCCL all-reduce: every
    rank stores its values straight into every peer's mailbox (torch symmetric memory provides the peer mappings) and
    waits for the peers' sequence flags.  An ordinary kernel launch on the current stream: capturable in CUDA graphs,
    a few microseconds of latency, deterministic (rank-ordered sum).  Needs P2P access between the GPUs of the group.
    """

    def __init__(self, n_values, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        assert 1 <= n_values <= 8
        self.n = n_values
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._lib = ops.lib()
        nfl = self._lib.b200rl_p2p_mailbox_floats(self.world)
        self.mailbox = symm_mem.empty(nfl, dtype=torch.float32, device=self.device)
        self.mailbox.zero_()
        self.handle = symm_mem.rendezvous(self.mailbox, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.seq = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.buf = torch.zeros(n_values, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # every mailbox is zeroed and mapped before the first exchange

    def reduce(self, src):
        """src: device tensor with at least n float32 values (e.g. the kernel's raw loss vector). Returns self.buf."""
        from . import _lib, ops
        rc = self._lib.b200rl_p2p_allreduce_mean(src.data_ptr(), self.ptrs.data_ptr(), self.rank, self.world, self.n,
                                                 self.seq.data_ptr(), self.buf.data_ptr(), ops.stream_ptr())
        _lib.check(rc, 'b200rl_p2p_allreduce_mean')
        return self.buf


class FusedLossExchange:
    """State of the exchange that rides on the launches of the one-launch learner step (``b200rl_gae_ppo_fwd_grad_dp``,
    csrc/common.cuh): step q's loss-finalisation launch stages ``{q, value}`` locally; step q+1's streaming kernel -- six
    warps of its first CTA, while they wait for their first chunk anyway -- consumes the entries of step q-1 of all ranks
    and publishes the staged word into every peer's mailbox over NVLink.  No collective call, no extra launch, no forked graph
    branch, and neither the remote stores nor the mailbox reads on the critical path.
    ``out_mean[:6]`` holds the mean over ranks of the latest consumed step (two steps behind the local losses),
    ``out_mean[8:14]`` the step before; ``drain()`` after the last step delivers the last step's mean (and the one before).

    Same contract as ``LossAllReduce`` (mean of equal-sized rank means, ding/utils/pytorch_ddp_dist_helper.py:38-47);
    every rank must launch the same sequence of steps.  Needs P2P access between the GPUs (torch symmetric memory).
    """

    def __init__(self, device, group=None, n_values=6):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.n = n_values
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._lib = ops.lib()
        nfl = self._lib.b200rl_p2p_mailbox_floats(self.world)
        self.mailbox = symm_mem.empty(nfl, dtype=torch.float32, device=self.device)
        self.mailbox.zero_()
        self.handle = symm_mem.rendezvous(self.mailbox, self.group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.seq = torch.zeros(24, dtype=torch.int32, device=self.device)  # staged tags, staged values, consumed tags
        self.out_mean = torch.zeros(16, dtype=torch.float32, device=self.device)  # [0:8] latest mean, [8:16] the one before
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # every mailbox is zeroed and mapped before the first exchange

    def args(self):
        """(mailbox_ptrs_dev, rank, world, seq_dev, out_mean) for ``b200rl_gae_ppo_fwd_grad_dp``."""
        return self.ptrs.data_ptr(), self.rank, self.world, self.seq.data_ptr(), self.out_mean.data_ptr()

    def drain(self):
        """Consume the last launched step's entries (one tiny kernel on the current stream). Returns ``out_mean``."""
        from . import _lib, ops
        rc = self._lib.b200rl_p2p_drain_mean(self.ptrs.data_ptr(), self.rank, self.world, self.n, self.seq.data_ptr(),
                                             self.out_mean.data_ptr(), ops.stream_ptr())
        _lib.check(rc, 'b200rl_p2p_drain_mean')
        return self.out_mean

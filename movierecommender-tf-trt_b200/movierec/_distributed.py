"""Data-parallel training of one NeuMF model over the GPUs of a box.

One process per GPU (`torch.distributed`, NCCL over NVLink 5 / NVSwitch on GPUs, gloo in CPU
tests).  Tables are replicated.  A global batch is split across ranks BY WHOLE GROUPS (a group = one
positive and its negatives, so the ranking metrics stay rank-local); every rank runs the fused
forward/backward on its groups with the gradient scale 1/B_global, the flat gradient buffer
(dense-layer gradients + the embedding gradient tables) is summed with ONE all-reduce, and every
rank applies the identical optimizer update.  The result equals the single-GPU step on the global
batch up to summation order.  The reference is single-process: there is no reference
collective to mirror; this is the exchange step the data-parallel path needs and nothing more.
"""

import torch
import torch.distributed as dist


def split_groups(num_groups, world_size, rank):
    """[lo, hi) of the groups rank `rank` owns: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(num_groups), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(users, items, labels, group, world_size, rank):
    """Rank-local slice of a global batch laid out as the generator does (groups contiguous)."""
    n = len(labels)
    if n % group:
        raise ValueError("batch of {} rows is not divisible by the group width {}".format(n, group))
    lo, hi = split_groups(n // group, world_size, rank)
    sl = slice(lo * group, hi * group)
    return users[sl], items[sl], labels[sl]


class DataParallelNeuMF(object):
    """Wraps a NeuMFEngine replica; `train_step` takes the RANK-LOCAL rows of a global batch."""

    def __init__(self, engine, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if engine.table_mode != "dense":
            raise NotImplementedError("replicated data parallelism all-reduces dense gradient tables; "
                                      "use table_mode='dense'")
        self.engine = engine
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights."""
        e = self.engine
        for t in [e.dense] + list(e._tables.values()):
            dist.broadcast(t, src, group=self.group)

    def train_step(self, users, items, labels, global_rows, group=0, k=0):
        """Local forward/backward -> all-reduce(sum) of the flat gradients -> identical update.
        Returns the rank-local step outputs (loss/hit/dcg sums over the local rows)."""
        e = self.engine
        out = e.train_grads(users, items, labels, group=group, k=k, inv_global_batch=1.0 / float(global_rows))
        for t in e.gradient_tensors():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        e.apply()
        return out

    def all_reduce_sums(self, t):
        """Sum metric / loss accumulators over ranks (reporting only)."""
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

The reference has no distributed code (`param_search.sh:79-95` only launches independent replicas).  Training shards
naturally by example: every rank holds the full weights and its slice of the batch, and because the loss is a SUM over
examples (`nn_utils/core.py:265-267`) the only exchange is one all-reduce(SUM) of the flat gradient buffer before the
global-norm clip + Adam update (`core.py:94-103`) -- N ranks x B/N examples then reproduce the single-GPU batch-B step.
With `weighted_classes` (as executed: mean CE, core.py:244-267) each rank scales by 1/B_local, so the reduced gradient
must additionally be divided by the world size.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous [begin, end) slice of n examples for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_ids(ids, rank, world, group_key=None):
    """Slice an id list across ranks.  With `group_key` (e.g. the image of a mention|box pair, icl_affinity_lstm.py:36-40)
    whole groups stay on one rank so a rank's box rows are local."""
    ids = list(ids)
    if group_key is None:
        b, e = shard_range(len(ids), rank, world)
        return ids[b:e]
    groups, order = {}, []
    for i in ids:
        k = group_key(i)
        if k not in groups:
            groups[k] = []
            order.append(k)
        groups[k].append(i)
    b, e = shard_range(len(order), rank, world)
    return [i for k in order[b:e] for i in groups[k]]


def global_offsets(local_counts, rank):
    """Global id of this rank's first sequence / example: dropout masks are keyed on global ids, so the result does not
    depend on how many ranks share the batch."""
    return int(np.sum(local_counts[:rank]))


def allreduce_sum_(tensor, group=None):
    """In-place SUM all-reduce of the flat gradient buffer (a torch tensor on the rank's device)."""
    import torch.distributed as td
    td.all_reduce(tensor, op=td.ReduceOp.SUM, group=group)
    return tensor

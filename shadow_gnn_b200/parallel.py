"""Data parallelism over target nodes: independent units, one gradient exchange per step.

Every rank samples and trains on its own slice of the epoch's target order; graph, PPR tables and features are replicated.
The only collective is a SUM all-reduce of the flat fp32 gradient buffer (NCCL on the GPUs; the same code runs on gloo for
the CPU tests); the optimizer then scales by 1/world so that the gradient-norm clip (models.py:223 of the reference) is
applied to the averaged gradient, exactly as in the single-GPU run.
"""
import numpy as np
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_targets(targets, rank, world, batch_size=1):
    """Strided slice of the epoch's target order, truncated so that every rank owns the same number of whole batches
    (all ranks must run the same number of steps, or the all-reduce would hang)."""
    targets = np.asarray(targets)
    per_rank = (targets.size // world // batch_size) * batch_size
    if per_rank == 0:
        per_rank = targets.size // world
    return targets[rank::world][:per_rank]


def allreduce_flat_gradients(flat_grad):
    """one collective per step over the whole gradient bucket; returns the factor the optimizer has to apply (1/world)"""
    _, world = world_info()
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world

"""Data parallelism over target nodes: independent units, one gradient exchange per step.

Every rank samples and trains on its own slice of the epoch's target order; graph, PPR tables and features are replicated.
The only collective is a SUM all-reduce of the flat fp32 gradient buffer (NCCL on the GPUs; the same code runs on gloo for
the CPU tests); the optimizer then scales by 1/world so that the gradient-norm clip (models.py:223 of the reference) is
applied to the averaged gradient, exactly as in the single-GPU run.
"""
import numpy as np
import torch
import torch.distributed as dist


def _join_wgrad():
    from .ops import join_wgrad          # (ops imports the CUDA library: resolved at call time so the gloo / CPU tests can import this module)
    join_wgrad()


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_targets(targets, rank, world, batch_size=1):
    """Strided slice of the epoch's target order.  Every rank must run the same number of steps (or the all-reduce would hang), so the
    slices are PADDED to the same whole number of batches by wrapping around to the head of the rank's own slice: no
    target is dropped, at most `batch_size * world - 1` are seen twice in an epoch."""
    targets = np.asarray(targets)
    if targets.size == 0 or world <= 1:
        return targets                                                   # one rank: the reference's own epoch (short last batch, minibatch.py:252-262)
    mine = targets[rank::world]
    longest = -(-targets.size // world)                                  # ceil: the longest strided slice
    per_rank = -(-longest // batch_size) * batch_size
    if mine.size == 0:                                                   # fewer targets than ranks: borrow from the global order
        mine = targets[:1]
    reps = -(-per_rank // mine.size)
    return np.concatenate([mine] * reps)[:per_rank] if mine.size < per_rank else mine[:per_rank]


def allreduce_flat_gradients(flat_grad):
    """one collective per step over the whole gradient bucket; returns the factor the optimizer has to apply (1/world)"""
    _, world = world_info()
    if world > 1:
        _join_wgrad()
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def allreduce_bucket(flat_grad, lo, hi, side_stream):
    """inside a stream capture: the bucket flat_grad[lo:hi] (layers whose backward is already done) is reduced on `side_stream`
    while the rest of the backward pass runs on the current stream; returns nothing -- finish with `join_buckets`"""
    cur = torch.cuda.current_stream()
    side_stream.wait_stream(cur)
    from .ops import wgrad_stream
    ws = wgrad_stream(flat_grad.device)                    # weight gradients of the tail layers run on ops' side stream: the exchange waits
    if ws is not None:                                     # for them there, the backward pass on the current stream does not
        side_stream.wait_stream(ws)
    with torch.cuda.stream(side_stream):
        dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM)


def join_buckets(flat_grad, split, side_stream):
    """the head bucket on the current stream, then wait for the tail bucket"""
    _join_wgrad()
    if split > 0:
        dist.all_reduce(flat_grad[:split], op=dist.ReduceOp.SUM)
    torch.cuda.current_stream().wait_stream(side_stream)

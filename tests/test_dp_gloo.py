"""CPU tests of the multi-process path (gloo, world_size 2): target partitioning and the single gradient exchange.
DP(2) on two half batches must give the gradient DP(1) computes on the whole batch, before and after the norm clip."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shadow_gnn_b200.parallel import allreduce_flat_gradients, partition_targets


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))


def _flat_grad(model, x, y):
    model.zero_grad()
    torch.nn.functional.cross_entropy(model(x), y).backward()
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(64, 8, generator=g), torch.randint(0, 3, (64,), generator=g)
    targets = np.arange(64)
    mine = partition_targets(targets, rank, world, batch_size=8)
    flat = _flat_grad(_model(), X[mine], Y[mine])
    scale = allreduce_flat_gradients(flat)
    torch.save(dict(grad=flat * scale, mine=mine), os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_dp2_equals_dp1(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt", weights_only=False) for i in range(world)]
    assert torch.equal(r[0]["grad"], r[1]["grad"])                              # both ranks hold the same reduced gradient
    assert sorted(np.concatenate([r[0]["mine"], r[1]["mine"]]).tolist()) == list(range(64))     # disjoint cover, equal shares
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(64, 8, generator=g), torch.randint(0, 3, (64,), generator=g)
    want = _flat_grad(_model(), X, Y)                                           # DP(1): the whole global batch
    assert torch.allclose(r[0]["grad"], want, rtol=1e-5, atol=1e-6)
    clip = lambda v: v * min(1.0, 5.0 / (float(v.norm()) + 1e-6))               # clip_grad_norm_ on the REDUCED gradient
    assert torch.allclose(clip(r[0]["grad"] * 100), clip(want * 100), rtol=1e-5, atol=1e-6)


def test_partition_targets_equal_whole_batches():
    """every rank gets the same whole number of batches; the tail of the epoch is padded by wrap-around, never dropped"""
    t = np.arange(1000)
    parts = [partition_targets(t, r, 8, batch_size=32) for r in range(8)]
    assert len({p.size for p in parts}) == 1 and parts[0].size % 32 == 0 and parts[0].size == 128
    assert set(np.concatenate(parts).tolist()) == set(range(1000))              # nothing dropped
    for r, p in enumerate(parts):
        assert set(p.tolist()) == set(t[r::8].tolist())                          # padding comes from the rank's own slice
    assert partition_targets(t, 0, 1, 32).size == 1000                           # one rank: the reference's epoch, short last batch included
    assert partition_targets(np.arange(3), 5, 8, 2).size == 2                    # fewer targets than ranks

"""2+ GPU check of the peer-memory gradient exchange (ops._P2PGrad + shadow_p2p_adam_clip_step_f32) against NCCL all-reduce + the local optimizer
step, eager and inside a CUDA graph.  Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_p2p.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from shadow_gnn_b200.ops import FlatAdamClip
    dev = torch.device("cuda")
    shapes = [(256, 100), (256,), (256, 256), (2, 256), (47, 256), (47,), (3,)]

    def make(p2p):
        os.environ["SHADOW_P2P"] = "1" if p2p else "0"
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in shapes]
        return params, FlatAdamClip(params, lr=0.01, max_norm=5.0)
    pa, oa = make(True)
    pb, ob = make(False)
    assert oa.p2p is not None and ob.p2p is None
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)

    def one_step(graphs=None):
        oa.zero_grad(); ob.zero_grad()
        fill = torch.randn(ob.grad.numel(), device=dev, generator=g) * (3.0 if rank == 0 else 0.5)
        oa.grad.copy_(fill); ob.grad.copy_(fill)
        oa.step()
        dist.all_reduce(ob.grad)
        ob.step(1.0 / world)
    worst = 0.0
    for it in range(6):
        one_step()
        torch.cuda.synchronize()
        d = float((oa.flat - ob.flat).abs().max())
        worst = max(worst, d)
        assert not oa.p2p.error()
    # the same exchange replayed from a CUDA graph (what GraphedTrainer captures): zero -> fill -> step
    src = torch.zeros_like(ob.grad)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        oa.zero_grad(); oa.grad.copy_(src); oa.step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    ob.zero_grad(); ob.grad.copy_(src); dist.all_reduce(ob.grad); ob.step(1.0 / world)
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        oa.zero_grad(); oa.grad.copy_(src); oa.step()
    for it in range(20):
        src.copy_(torch.randn(src.numel(), device=dev, generator=g))
        graph.replay()
        ob.zero_grad(); ob.grad.copy_(src); dist.all_reduce(ob.grad); ob.step(1.0 / world)
    torch.cuda.synchronize()
    d2 = float((oa.flat - ob.flat).abs().max())
    # replicas must hold identical weights (every rank sums the ranks' gradients in the same order)
    ref = oa.flat.clone(); dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, oa.flat)) if False else float((ref - oa.flat).abs().max())
    # timing: exchange + optimizer, p2p vs NCCL
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    gn = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gn, capture_error_mode="thread_local"):
        ob.zero_grad(); ob.grad.copy_(src); dist.all_reduce(ob.grad); ob.step(1.0 / world)
    for gr in (graph, gn):
        for _ in range(5): gr.replay()
    torch.cuda.synchronize(); dist.barrier()
    ev[0].record()
    for _ in range(50): graph.replay()
    ev[1].record(); torch.cuda.synchronize(); dist.barrier()
    ev[2].record()
    for _ in range(50): gn.replay()
    ev[3].record(); torch.cuda.synchronize()
    print(f"rank {rank}/{world}: max |p2p - nccl| eager {worst:.3e}, graphed {d2:.3e}; replica drift {same:.3e}; error flag {oa.p2p.error()}; "
          f"zero+fill+exchange+adam per step: p2p {ev[0].elapsed_time(ev[1]) / 50 * 1e3:.1f} us, nccl {ev[2].elapsed_time(ev[3]) / 50 * 1e3:.1f} us", flush=True)
    assert worst < 1e-4 and d2 < 1e-4 and not oa.p2p.error()      # fp32 sums of `world` gradients in a different order than NCCL's
    assert same == 0.0, "replicas must not drift: the norm is summed in a fixed order on every rank"
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

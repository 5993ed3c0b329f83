"""Kernel-time breakdown of the training step (torch.profiler / CUPTI, eager steps so every kernel is visible)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from shadow_gnn_b200 import minibatch as MB
from shadow_gnn_b200.models import DeepGNN
from shadow_gnn_b200.train import GraphedTrainer
sys.argv = ["bench.py", "--task", "train"] + os.environ.get("BENCH_ARGS", "").split()
args = bench.parse()
ctx = bench.Ctx(args)
dev, F, C, B = ctx.dev, ctx.F, ctx.C, args.batch
labels = torch.from_numpy(np.random.default_rng(7).integers(0, C, ctx.N)).to(dev)
share = ctx.share.numpy()[:8192]
cfg = {"batch_size": B, "configs": [{"method": "ppr", "k": [150], "threshold": [0.0], "epsilon": [1e-5]}]}
adjs = {m: (ctx.g["indptr"], ctx.g["indices"]) for m in range(3)}
mb = MB.MinibatchShallowExtractor("g", None, adjs, {0: share, 1: share[:B], 2: share[:B]}, cfg, set(), None, ctx.feat, labels, F, True, 1, seed_cpp=1, num_subg_per_batch=4096)
model = DeepGNN(F, F, C, 0, bench.ARCH, [], 1, dict(dropout=bench.TRAIN_CFG["dropout"], dropedge=bench.TRAIN_CFG["dropedge"], lr=bench.TRAIN_CFG["lr"], ensemble_dropout="none"), "node").to(dev)
mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
for _ in range(5):
    model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total if hasattr(e, "self_device_time_total") else e.self_cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per step: {tot/10:.1f} us")
for k, t, c in rows[:28]:
    print(f"{t/10:9.1f} us/step {c/10:6.1f}x {t/tot*100:5.1f}%  {k[:100]}")
# graphed step timing + static-load overhead
tr = GraphedTrainer(model, mb, row_cap=B * 151, edge_cap=B * 151 * 64)
for _ in range(10): tr.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): tr.step()
torch.cuda.synchronize()
print("graphed step wall ms:", (time.perf_counter() - t0) / 50 * 1e3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(50): tr.graph.replay()
ev[1].record(); torch.cuda.synchronize()
print("graph replay only ms:", ev[0].elapsed_time(ev[1]) / 50)
with profile(activities=[ProfilerActivity.CUDA]) as prof2:
    for _ in range(10):
        tr.step()
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total, e.count) for e in prof2.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"GRAPHED: total kernel time per step: {tot/10:.1f} us, kernels per step: {sum(r[2] for r in rows)/10:.0f}")
for k, t, c in rows[:22]:
    print(f"G {t/10:9.1f} us/step {c/10:6.1f}x avg {t/max(c,1):7.2f}  {k[:110]}")

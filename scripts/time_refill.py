"""Scratch: wall time of one super-batch refill (MinibatchShallowExtractor.par_graph_sample + what the graph trainer needs from it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from shadow_gnn_b200 import minibatch as MB, ParallelSampler as PS
sys.argv = ["bench.py", "--task", "train"]
args = bench.parse()
ctx = bench.Ctx(args)
dev, F, C, B = ctx.dev, ctx.F, ctx.C, 32
labels = torch.from_numpy(np.random.default_rng(7).integers(0, C, ctx.N)).to(dev)
share = ctx.share.numpy()
cfg = {"batch_size": B, "configs": [{"method": "ppr", "k": [150], "threshold": [0.0], "epsilon": [1e-5]}]}
adjs = {m: (ctx.g["indptr"], ctx.g["indices"]) for m in range(3)}
for sb in (160, 1600, 4096):
    mb = MB.MinibatchShallowExtractor("g", None, adjs, {0: share, 1: share[:B], 2: share[:B]}, cfg, set(), None, ctx.feat, labels, F, True, 1, seed_cpp=1, num_subg_per_batch=sb)
    mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
    def T():
        torch.cuda.synchronize(); return time.perf_counter()
    for it in range(6):
        mb.pool[MB.TRAIN][0].clear()
        t0 = T()
        s = mb.graph_sampler[MB.TRAIN]
        batches = s.sample_to_device(mb.sampler_cfgs[MB.TRAIN], [set()])
        t1 = T()
        feat = PS.gather_rows(mb.feat_full, batches[0].orig_node)
        t2 = T()
        sbo = MB._SuperBatch(batches[0], feat, {}, 1)
        t3 = T()
        sbo.take_canonical(B)
        t4 = T()
        if it >= 2:
            print(f"sb {sb}: sample {1e3*(t1-t0):.2f} ms, gather {1e3*(t2-t1):.2f}, superbatch ctor {1e3*(t3-t2):.2f}, canonical+first take {1e3*(t4-t3):.2f}, total {1e3*(t4-t0):.2f}", flush=True)

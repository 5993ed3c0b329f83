"""Exploration (scratch, CPU only): expected shared-memory wavefronts per warp-wide membership probe for candidate hash-table layouts of
the PPR fast path, on the real key stream of the S-products stand-in (node sets = PPR top-150 rows, probe keys = the scanned slots).

Model: 32 banks x 4 B; a request wider than 4 B per lane is served in passes (8 B: two half-warps, 16 B: four quarter-warps); within a
pass the cost is the largest number of DISTINCT addresses that fall into one bank (same address = broadcast).  Prints wavefronts per
32 probed slots and the overflow statistics of each layout; compare with ncu's l1tex__data_pipe_lsu_wavefronts_mem_shared
(profiles/r1_ppr_induce_warp_kernel.md: 4.3 k wavefronts per subgraph with 16-byte probes, 2.9 k with 8-byte probes).

    python scripts/model_probe_wavefronts.py /tmp/work      # directory holding indptr.npy, indices.npy, nb.npy, ln.npy, t.npy (scripts/deg_stats-style dump)
"""
import sys

import numpy as np

GOLD = np.uint64(2654435761)


def hash_bucket(keys, nb):
    lg = int(np.log2(nb))
    return ((keys.astype(np.uint64) * GOLD) & np.uint64(0xFFFFFFFF)) >> np.uint64(32 - lg)


def wavefronts(addr_bytes, width):
    """addr_bytes: [n_warps, 32] byte address of each lane's probe; width in bytes (4, 8, 16)"""
    lanes_per_pass = {4: 32, 8: 16, 16: 8}[width]
    total = 0
    for p0 in range(0, 32, lanes_per_pass):
        a = addr_bytes[:, p0:p0 + lanes_per_pass]
        word = a // 4                                            # first 4-byte word; a width-W access covers W/4 consecutive banks
        cost = np.zeros(a.shape[0], dtype=np.int64)
        for w in range(a.shape[0]):
            uniq = np.unique(word[w])
            banks = (uniq % 32)
            cost[w] = np.bincount(banks, minlength=32).max()
        total += cost.sum()
    return total


def main(d):
    indptr, indices = np.load(f"{d}/indptr.npy"), np.load(f"{d}/indices.npy")
    nb, ln, t = np.load(f"{d}/nb.npy"), np.load(f"{d}/ln.npy"), np.load(f"{d}/t.npy")
    S = min(64, t.size)
    layouts = [("1-key x 1024 (4 KB, 4-byte probe)", 1, 1024, 4), ("2-key x 512 (4 KB, 8-byte probe)", 2, 512, 8),
               ("2-key x 1024 (8 KB, 8-byte probe)", 2, 1024, 8), ("4-key x 256 (4 KB, 16-byte probe)", 4, 256, 16)]
    res = {name: dict(wf=0, probes=0, ovf=[]) for name, *_ in layouts}
    for i in range(S):
        nodes = np.unique(np.concatenate([nb[i, :ln[i]], [t[i]]])).astype(np.uint32)
        slots = np.concatenate([indices[indptr[v]:indptr[v + 1]] for v in nodes]).astype(np.uint32)
        pad = (-slots.size) % 32
        keys = np.concatenate([slots, np.zeros(pad, np.uint32)]).reshape(-1, 32)
        for name, bk, nbuckets, width in layouts:
            b = hash_bucket(nodes, nbuckets)
            cnt = np.bincount(b.astype(np.int64), minlength=nbuckets)
            res[name]["ovf"].append(int(np.maximum(cnt - bk, 0).sum()))
            addr = hash_bucket(keys.reshape(-1), nbuckets).astype(np.int64).reshape(-1, 32) * (4 * bk)
            res[name]["wf"] += wavefronts(addr, width)
            res[name]["probes"] += keys.shape[0]
    for name, *_ in layouts:
        r = res[name]
        o = np.array(r["ovf"])
        print(f"{name:36s} wavefronts per warp-wide probe {r['wf'] / r['probes']:5.2f}   overflow keys per subgraph: mean {o.mean():.2f}, "
              f"P(>0) {np.mean(o > 0):.2f}, P(>3) {np.mean(o > 3):.2f}, max {o.max()}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/work")

"""Shared helpers for the parity tests: golden-fixture access and subgraph comparison."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_golden.npz")
FIELDS = ("indptr", "indices", "node", "edge_index", "target", "hop", "ppr", "drnl")


class Golden:
    def __init__(self):
        self.z = np.load(GOLDEN)
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.indptr, self.indices = self.z["indptr"], self.z["indices"]
        self.ppr_ptr, self.ppr_neighs, self.ppr_scores = self.z["ppr_ptr"], self.z["ppr_neighs"], self.z["ppr_scores"]

    @property
    def num_cases(self):
        return len(self.meta["cases"])

    def case(self, ci):
        c = self.meta["cases"][ci]
        subgs = None
        for f in FIELDS:
            lens = self.z[f"c{ci}_{f}_len"]
            flat = self.z[f"c{ci}_{f}"]
            off = np.concatenate([[0], np.cumsum(lens)])
            parts = [flat[off[i]:off[i + 1]] for i in range(lens.size)]
            if subgs is None:
                subgs = [dict() for _ in parts]
            for s, p in zip(subgs, parts):
                s[f] = p
        return c["cfg"], tuple(c["aug"]), self.z[f"c{ci}_targets"], c["calls"], subgs


def assert_subgraph_equal(a, b, ctx=""):
    """Bit-exact: integers by value, ppr by bit pattern."""
    for f in FIELDS:
        x, y = np.asarray(a[f]), np.asarray(b[f])
        assert x.size == y.size, f"{ctx} field {f}: size {x.size} vs {y.size}"
        if f == "ppr":
            assert x.astype(np.float32).tobytes() == y.astype(np.float32).tobytes(), f"{ctx} field ppr differs"
        else:
            assert np.array_equal(x.astype(np.int64), y.astype(np.int64)), f"{ctx} field {f} differs:\n{x}\n{y}"


def det_fill(module):
    """Deterministic, platform-independent parameter values keyed by parameter NAME (numpy Generator streams), so that the reference model in
    tests/golden/make_r2_golden.py and our model on the GPU box hold identical weights without shipping a multi-MB state_dict."""
    import zlib
    import torch
    with torch.no_grad():
        for name, p in module.named_parameters():
            a = np.random.default_rng(zlib.crc32(name.encode())).standard_normal(tuple(p.shape), dtype=np.float32)
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "weight":
                a = a / np.float32(np.sqrt(p.shape[-1]))
            elif leaf == "scale":
                a = np.float32(1) + np.float32(0.3) * a
            elif leaf in ("attention", "q"):
                a = np.float32(0.3) * a
            else:                                   # bias, offset, eps
                a = np.float32(0.1) * a
            p.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(p.device))
    return module


def grad_signature(named_grads):
    """per-tensor (L2 norm, projection on a name-keyed random direction) + the full tensor when it is small: pins a gradient without shipping it"""
    import zlib
    out = {}
    for name, g in named_grads:
        g = np.asarray(g, np.float64).ravel()
        d = np.random.default_rng(zlib.crc32(name.encode()) ^ 0x5bd1e995).standard_normal(g.size)
        out[f"{name}|norm"] = np.float64(np.sqrt((g * g).sum()))
        out[f"{name}|proj"] = np.float64((g * d).sum() / np.sqrt(g.size))
        if g.size <= 512:
            out[f"{name}|full"] = g.astype(np.float32)
    return out

"""Import the UNMODIFIED reference Python stack (shaDow.layers / shaDow.models / graph_engine.frontend) in this container.

Only used to GENERATE golden fixtures (make_layer_golden.py) and by the optional live-reference tests; it needs
/root/reference, which does not exist on the GPU box.  What it does:
  * stubs the three third-party modules that are not installed (torch_scatter, torch_geometric, ogb) with minimal
    equivalents written against their public semantics;
  * loads graph_engine/frontend/graph.py from the reference tree with its two mutable dataclass defaults rewritten to
    default_factory (Python >= 3.11 rejects them) -- in memory only, the reference tree is never modified;
  * imports with cwd = reference root and a fake argv, because shaDow/globals.py parses argv and reads CONFIG_TEMPLATE.yml at import.
"""
import importlib.util
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "shaDow"))


def _stub_modules():
    import torch
    ts = types.ModuleType("torch_scatter")

    def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
        assert src.dim() == 1 and index.dim() == 1
        n = int(index.max()) + 1 if dim_size is None else dim_size
        red = {"sum": "sum", "add": "sum", "max": "amax", "min": "amin", "mean": "mean"}[reduce]
        init = torch.zeros(n, dtype=src.dtype, device=src.device)
        return init.scatter_reduce(0, index, src, reduce=red, include_self=False)
    ts.scatter = scatter
    sys.modules["torch_scatter"] = ts

    tg = types.ModuleType("torch_geometric"); tgn = types.ModuleType("torch_geometric.nn"); tgu = types.ModuleType("torch_geometric.utils")

    def global_sort_pool(x, batch, k):
        # published PyG semantics: per graph, sort nodes by the last channel (descending), keep k rows, zero-pad, flatten
        B = int(batch.max()) + 1
        out = []
        for b in range(B):
            xb = x[batch == b]
            order = torch.argsort(xb[:, -1], descending=True)
            xb = xb[order][:k]
            if xb.shape[0] < k:
                xb = torch.cat([xb, xb.new_zeros(k - xb.shape[0], xb.shape[1])], 0)
            out.append(xb.reshape(-1))
        return torch.stack(out, 0)
    tgn.global_sort_pool = global_sort_pool
    for n in ("negative_sampling", "add_self_loops", "to_undirected"):
        setattr(tgu, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("link-task utilities are not stubbed")))
    tg.nn, tg.utils = tgn, tgu
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgn, "torch_geometric.utils": tgu})

    for name in ("ogb", "ogb.nodeproppred", "ogb.linkproppred"):
        m = types.ModuleType(name)
        m.Evaluator = type("Evaluator", (), {"__init__": lambda self, *a, **k: None})
        sys.modules[name] = m


_loaded = {}


def load():
    """returns dict(layers=<module shaDow.layers>, models=<module shaDow.models>, graph_utils=..., graph=...)"""
    if _loaded:
        return _loaded
    assert available(), "/root/reference is not present"
    _stub_modules()
    for p in (REF, os.path.join(REF, "para_graph_sampler")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the compiled reference sampler must be importable as `ParallelSampler`
    from oracle import oracle as O
    ps = O.load_ref()
    if ps is not None:
        sys.modules["ParallelSampler"] = ps
    cwd, argv = os.getcwd(), sys.argv
    os.chdir(REF)
    sys.argv = ["shim", "--dataset", "arxiv", "--no_log", "--gpu", "-1"]
    try:
        src = open(os.path.join(REF, "para_graph_sampler/graph_engine/frontend/graph.py")).read()
        src = src.replace("from dataclasses import dataclass, InitVar", "from dataclasses import dataclass, InitVar, field")
        for f in ("hop", "ppr", "drnl", "node", "edge_index", "target"):
            src = src.replace(f"    {f:<16}: np.ndarray = np.array([])", f"    {f:<16}: np.ndarray = field(default_factory=lambda: np.array([]))")
        spec = importlib.util.spec_from_loader("graph_engine.frontend.graph", loader=None)
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = "<patched in memory>"
        sys.modules["graph_engine.frontend.graph"] = mod
        exec(compile(src, "graph_engine/frontend/graph.py", "exec"), mod.__dict__)
        import graph_engine.frontend  # noqa: F401  (its __init__ picks the patched graph module up from sys.modules)
        import graph_engine.frontend.graph_utils as graph_utils
        import shaDow.layers as layers
        import shaDow.models as models
    finally:
        os.chdir(cwd)
        sys.argv = argv
    _loaded.update(layers=layers, models=models, graph_utils=graph_utils, graph=mod)
    return _loaded

#!/usr/bin/env python3
"""Measurements of the other BASELINE.json configurations (C1, C2, C3, C5); `bench.py` holds the
headline C4 contract.  One JSON line per configuration.

  python bench_configs.py C1 C2 C3          real datasets (latency-bound): us per full / dirty-path evaluation,
                                            MCMC generations/s, batched NNI scoring, the 1-core reference beside it
  python bench_configs.py C5 [--patterns N] synthetic 512 taxa x N x 64 states GTR+G4 (FP64 tensor path)
"""
import argparse
import io
import json
import os
import statistics
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

DATA = os.path.join(REPO, "tests", "golden", "data")
REAL = {
    "C1": ("narrow.phy", "readBinaryPhy", "bin", "F81"),
    "C2": ("IELex-2016.prog.phy", "readPhy", "multi", "JC"),
    "C3": ("ielex_multistate.phy", "readPhy", "multi", "F81"),
}


def _reference_times(fname, reader, dtype, model, n_dirty=20):
    """(ms full matML, ms cache_matML, lnL) of the compiled reference, 1 core, in a fresh interpreter."""
    code = f'''
import sys, io, time, random, contextlib, json
import numpy as np
sys.path.insert(0, {os.path.join(REPO, "oracle", "_ref")!r})
import config, utils, mcmc_gamma, ML_gamma
np.random.seed(1234); random.seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    (config.N_TAXA, config.N_CHARS, config.ALPHABET, sd, config.LEAF_LLMAT, config.TAXA, config.N_SITES) = utils.{reader}({os.path.join(DATA, fname)!r})
config.IN_DTYPE, config.MODEL, config.N_NODES = {dtype!r}, {model!r}, 2 * config.N_TAXA - 1
st = mcmc_gamma.state_init()
a = (config.N_SITES, config.N_TAXA, config.N_CATS)
lnl, cache = ML_gamma.matML(st["pi"], st["root"], config.LEAF_LLMAT, st["postorder"], st["transitionMat"], *a)
t = []
for _ in range(5):
    t0 = time.perf_counter(); ML_gamma.matML(st["pi"], st["root"], config.LEAF_LLMAT, st["postorder"], st["transitionMat"], *a); t.append(time.perf_counter() - t0)
rev = mcmc_gamma.adjlist2reverse_nodes_dict(st["tree"])
td = []
rng = random.Random(3)
for _ in range({n_dirty}):
    p, c = rng.choice(list(st["tree"]))
    path = mcmc_gamma.get_path2root(rev, c, st["root"])
    t0 = time.perf_counter(); ML_gamma.cache_matML(st["pi"], st["root"], config.LEAF_LLMAT, cache, path, st["postorder"], st["transitionMat"], *a); td.append(time.perf_counter() - t0)
print("@@" + json.dumps([sorted(t)[len(t)//2] * 1e3, sum(td) / len(td) * 1e3, float(lnl)]))
'''
    import subprocess
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("@@")]
    if not line:
        return None
    return json.loads(line[0][2:])


def run_real(name):
    import random
    from cybayes_b200 import config, likelihood
    from cybayes_b200.driver import load_alignment, run_chain
    from cybayes_b200.mcmc_gamma import (adjlist2nodes_dict, adjlist2reverse_nodes_dict, get_path2root, postorder,
                                         state_init)
    from cybayes_b200.ML_gamma import cache_matML, matML, score_proposals
    fname, reader, dtype, model = REAL[name]
    likelihood.reset_engines()
    np.random.seed(1234)
    random.seed(1234)
    load_alignment(os.path.join(DATA, fname), dtype, reader)
    config.MODEL = model
    st = state_init()
    a = (config.N_SITES, config.N_TAXA, config.N_CATS)
    lnl, cache = matML(st["pi"], st["root"], config.LEAF_LLMAT, st["postorder"], st["transitionMat"], *a)
    eng, _ = likelihood.engine_for(config.LEAF_LLMAT, config.N_CATS)
    wall, dev = [], []
    for _ in range(30):
        t0 = time.perf_counter()
        l2, c2 = matML(st["pi"], st["root"], config.LEAF_LLMAT, st["postorder"], st["transitionMat"], *a)
        wall.append(time.perf_counter() - t0)
        dev.append(eng.last_eval_ms())
        del c2
    parents = adjlist2reverse_nodes_dict(st["tree"])
    rng = random.Random(3)
    dwall, ddev, dlen = [], [], []
    for _ in range(40):
        p, c = rng.choice(list(st["tree"]))
        path = get_path2root(parents, c, st["root"])
        t0 = time.perf_counter()
        l3, c3 = cache_matML(st["pi"], st["root"], config.LEAF_LLMAT, cache, path, st["postorder"], st["transitionMat"], *a)
        dwall.append(time.perf_counter() - t0)
        ddev.append(eng.last_eval_ms())
        dlen.append(len(path))
        assert l3 == lnl
        del c3
    out = {"config": name, "dataset": fname, "model": model, "n_taxa": config.N_TAXA, "n_sites": config.N_SITES,
           "n_patterns": eng.n_patterns, "n_states": config.N_CHARS, "lnL": float(lnl),
           "full_eval": {"device_us": statistics.median(dev) * 1e3, "call_us": statistics.median(wall) * 1e6,
                         "evals_per_sec": 1.0 / statistics.median(wall)},
           "dirty_path": {"device_us": statistics.median(ddev) * 1e3, "call_us": statistics.median(dwall) * 1e6,
                          "mean_path_nodes": statistics.mean(dlen), "evals_per_sec": 1.0 / statistics.median(dwall)}}
    # batched NNI scoring: every NNI neighbour of the current tree in ONE launch vs one call each
    tree, root, N = st["tree"], st["root"], config.N_TAXA
    kids = adjlist2nodes_dict(tree)
    proposals = []
    for (aa, b) in list(tree):
        if b <= N:
            continue
        src = kids[aa][1] if kids[aa][0] == b else kids[aa][0]
        for tgt in kids[b]:
            t2 = dict(tree)
            sbl, tbl = t2.pop((aa, src)), t2.pop((b, tgt))
            t2[aa, tgt], t2[b, src] = tbl, sbl
            tm = []
            for k in range(config.N_CATS):
                tk = st["transitionMat"][k].copy()
                tk[aa, tgt], tk[b, src] = tk[b, tgt], tk[aa, src]
                tm.append(tk)
            order = postorder(adjlist2nodes_dict(t2), root)[::-1]
            dirty = [b] + get_path2root(adjlist2reverse_nodes_dict(t2), b, root)
            proposals.append((dirty, order, tm))
    t0 = time.perf_counter()
    batch = score_proposals(st["pi"], root, config.LEAF_LLMAT, cache, proposals)
    t_batch = time.perf_counter() - t0
    dev_batch = eng.last_eval_ms()
    t0 = time.perf_counter()
    single = [cache_matML(st["pi"], root, config.LEAF_LLMAT, cache, d, o, tm, *a)[0] for d, o, tm in proposals]
    t_single = time.perf_counter() - t0
    assert all(x == y for x, y in zip(batch.tolist(), single)), "batched scores differ from one-by-one"
    out["batched_nni"] = {"candidates": len(proposals), "device_us_one_launch": dev_batch * 1e3,
                          "call_ms_batched": t_batch * 1e3, "call_ms_one_by_one": t_single * 1e3,
                          "candidates_per_sec_device": len(proposals) / (dev_batch * 1e-3),
                          "check": "bit-identical to cache_matML per candidate"}
    ref = _reference_times(fname, reader, dtype, model)
    if ref:
        out["reference_1core"] = {"full_eval_ms": ref[0], "dirty_path_ms": ref[1], "lnL": ref[2],
                                  "lnL_rel_err": abs(float(lnl) - ref[2]) / abs(ref[2])}
        out["speedup_full"] = ref[0] * 1e-3 / statistics.median(wall)
        out["speedup_dirty"] = ref[1] * 1e-3 / statistics.median(dwall)
    likelihood.reset_engines()
    n_gen = 3000 if name == "C1" else 1000
    res = run_chain(os.path.join(DATA, fname), model, n_gen, 1000, dtype, os.path.join(tempfile.gettempdir(), "bc_" + name),
                    reader=reader, out=io.StringIO(), fast_spr=True)
    out["mcmc"] = {"gens_per_sec": res["gens_per_sec"], "generations": n_gen, "fast_spr": True}
    likelihood.reset_engines()
    print(json.dumps(out), flush=True)


def run_c5(n_patterns, n_taxa=512, S=64):
    """C5, on one GPU or -- under `python -m torch.distributed.run --nproc-per-node N bench_configs.py C5 ...` -- with the
    patterns sharded over N GPUs (one process each; the library all-reduces the scalar lnL over NCCL).  The full-size
    cache (209 GB) needs >= 2 GPUs."""
    from cybayes_b200 import _lib
    from cybayes_b200.engine import Engine
    from cybayes_b200.likelihood import _Plan
    from cybayes_b200.subst import gtr_eigensystem
    from cybayes_b200.synthetic import SyntheticAlignment, shard_bounds
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist  # host-side rendezvous only (gloo)
        dist.init_process_group("gloo")
        os.environ["CYBAYES_DEVICE"] = os.environ.get("LOCAL_RANK", "0")
    block_sites = min(n_patterns, 25000)
    aln = SyntheticAlignment(n_taxa, n_patterns, S, 20260102, block_sites=block_sites)
    lo, hi = shard_bounds(n_patterns, rank, world, block_sites) if world > 1 else (0, n_patterns)
    t0 = time.perf_counter()
    codes = aln.codes(lo, hi)
    t_gen = time.perf_counter() - t0
    n_local = hi - lo
    C = 4
    eng = Engine(codes, S, C)
    if world > 1:
        box = [eng.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)
    plan = _Plan(aln.edge_order())
    ekeys = list(aln.tree.keys())
    n_e = len(ekeys)
    block = eng.alloc_slots(n_e * C)
    slots = np.arange(block.base, block.base + n_e * C, dtype=np.int32)
    d = np.array([aln.tree[e] * r for r in aln.rates for e in ekeys])
    eng.mark(0)
    eng.queue_build(_lib.CB_MODEL_GTR_EIG, aln.pi, 0.0, gtr_eigensystem(aln.pi, aln.er), slots, d)
    eng.flush_builds()
    eng.mark(1)
    k1_ms = eng.mark_elapsed_ms()
    slot_of = {(k, e): block.base + k * n_e + i for k in range(C) for i, e in enumerate(ekeys)}
    pslots = np.array([[slot_of[k, e] for k in range(C)] for e in plan.edge_keys], dtype=np.int32)
    n_int_edges = sum(1 for (p, c) in ekeys if c > n_taxa)
    flops = C * n_patterns * (2.0 * S * S * n_int_edges + S * (n_taxa - 1) + 2.0 * S)
    alg_bytes = 16.0 * C * S * n_patterns * (n_taxa - 2) + 1.0 * n_taxa * n_patterns + 8.0 * n_patterns
    res = {}
    for label, snap in (("lnl_only", False), ("with_cache", True)):
        need = (n_taxa - 2) * C * S * n_local * 8 if snap else 0
        if need > 150e9:
            res[label] = {"skipped": f"cache of {need / 1e9:.0f} GB per GPU does not fit"}
            continue
        ms = []
        for _ in range(4):
            if dist is not None:
                dist.barrier()
            eng.mark(0)
            lnl, sn = eng.eval(None, plan.nodes, plan.children, pslots, aln.pi, want_snapshot=snap)
            eng.mark(1)
            ms.append(eng.mark_elapsed_ms() if world > 1 else eng.last_eval_ms())   # sharded: incl. the all-reduce
            if sn >= 0:
                eng.release_snapshot(sn)
        m = statistics.median(ms[1:])
        if dist is not None:
            import torch
            t = torch.tensor([m], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank
            m = float(t[0])
        res[label] = {"ms": m, "evals_per_sec": 1e3 / m, "tflops": flops / m / 1e9, "alg_GBps": alg_bytes / m / 1e6,
                      "lnL": lnl}
    if rank == 0:
        print(json.dumps({"config": "C5", "n_gpus": world, "n_taxa": n_taxa, "n_patterns": n_patterns, "n_states": S,
                          "model": "GTR", "patterns_per_gpu": n_local, "algorithmic_tflop": flops / 1e12,
                          "algorithmic_GB": alg_bytes / 1e9, "k1_pmat_build_ms": k1_ms, "n_matrices": n_e * C,
                          "data_generation_s": t_gen, **res}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--patterns", type=int, default=50000)
    a = ap.parse_args()
    for cfg in a.configs:
        if cfg in REAL:
            run_real(cfg)
        elif cfg == "C5":
            run_c5(a.patterns)
        else:
            raise SystemExit(f"unknown config {cfg}")

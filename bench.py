#!/usr/bin/env python3
"""Benchmark of the tree-likelihood hot path (contract in the task statement / DESIGN.md).

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA engine)
  python bench.py --impl reference [...]                      the reference's own CPU path

Workload (BASELINE.json, config C4): synthetic 1024 taxa x 1,000,000 binary site patterns,
GTR + discrete-Gamma-4, one *step* = one full Felsenstein pruning pass (log-likelihood
evaluation) keeping the partial cache.  For N > 1 the patterns are sharded over the ranks (one
process per GPU, torchrun), each rank runs the same op list on its slice and the only exchange is
the scalar NCCL all-reduce inside the library: strong scaling, `value` = whole-alignment evals/s.

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("CYBAYES_COMPRESS_MAX_SITES", "0")  # patterns are used as generated (weights 1)
REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

SEED = 20260101
BLOCK = 125000          # generation / sharding granule: 1M = 8 blocks
METRIC = "tree log-likelihood evals/sec"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--taxa", type=int, default=1024)
    ap.add_argument("--patterns", type=int, default=1000000)
    ap.add_argument("--sample-sites", type=int, default=20000, help="sites of the cpu_baseline sample")
    ap.add_argument("--ref-chunk", type=int, default=10000, help="sites per worker of --impl reference")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / dirty-path / MCMC extras")
    return ap.parse_args()


def workload_name(a):
    return (f"C4: synthetic {a.taxa} taxa x {a.patterns} binary site patterns, GTR + Gamma-4, "
            "full pruning pass (likelihood evaluation) with the partial cache kept")


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].strip() == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm
_REF = {}


def _ref_worker_init(taxa, patterns, chunk, counter):
    """Runs in a spawned process: import the compiled, unmodified reference; when `counter` is given,
    take the next block index and prepare that chunk of the alignment."""
    import random
    sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))
    import config as rconfig      # the reference's modules (oracle/_ref/*.so)
    import mcmc_gamma as rmcmc
    import ML_gamma as rml
    from cybayes_b200.synthetic import SyntheticAlignment
    aln = SyntheticAlignment(taxa, patterns, 2, SEED, block_sites=chunk)
    _REF.update(aln=aln, rml=rml, rconfig=rconfig, rmcmc=rmcmc, chunk=chunk)
    random.seed(1)
    if counter is not None:
        with counter.get_lock():
            idx = counter.value
            counter.value += 1
        _ref_prepare(idx % max(1, patterns // chunk))


def _ref_prepare(block_index):
    aln, rconfig, rmcmc = _REF["aln"], _REF["rconfig"], _REF["rmcmc"]
    codes = aln.codes(block_index * _REF["chunk"], (block_index + 1) * _REF["chunk"])
    eye = np.eye(2)
    rconfig.N_TAXA, rconfig.N_CHARS, rconfig.N_SITES = aln.n_taxa, 2, codes.shape[1]
    rconfig.MODEL, rconfig.IN_DTYPE = "GTR", "bin"
    _REF["leaves"] = {t + 1: np.ascontiguousarray(eye[codes[t]].T) for t in range(aln.n_taxa)}
    pi, er = aln.pi.copy(), aln.er.copy()
    _REF["tmats"] = [rmcmc.get_prob_t(pi, aln.tree, er, r) for r in aln.rates]
    _REF["edges"] = aln.edge_order()
    _REF["pi"] = pi
    return codes.shape[1]


def _ref_eval(_):
    aln = _REF["aln"]
    t0 = time.perf_counter()
    lnl, _cache = _REF["rml"].matML(_REF["pi"], aln.root, _REF["leaves"], _REF["edges"], _REF["tmats"],
                                    _REF["rconfig"].N_SITES, aln.n_taxa, 4)
    return time.perf_counter() - t0, float(lnl)


def reference_available():
    return os.path.exists(os.path.join(REPO, "oracle", "_ref")) and any(
        f.startswith("ML_gamma") and f.endswith(".so") for f in os.listdir(os.path.join(REPO, "oracle", "_ref")))


def run_reference_arm(a):
    """The reference's own Cython + NumPy matML on every host core (independent site chunks; the
    reference itself is single-threaded by construction, utils.pyx:3-7)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (oracle/build_ref.sh)"}))
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:  # each worker holds ~0.2 GB of leaves and produces a ~0.7 GB cache per call
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available / 1.5e9)))
    except Exception:
        pass
    ctx = mp.get_context("spawn")
    counter = ctx.Value("i", 0)
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(a.taxa, a.patterns, a.ref_chunk, counter)) as pool:
        for _ in range(a.warmup):
            pool.map(_ref_eval, range(cores), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_ref_eval, range(cores), chunksize=1)
        dt = (time.perf_counter() - t0) / a.steps
    sites_per_s = cores * a.ref_chunk / dt
    value = sites_per_s / a.patterns
    sample = (f"{cores} processes x {a.ref_chunk} sites of the same synthetic alignment per step; evals/s of the full "
              f"{a.patterns}-site alignment extrapolated linearly (sites are independent)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n_taxa": a.taxa, "n_patterns": a.patterns, "n_states": 2,
                   "n_cats": 4, "model": "GTR"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample(a, codes_sample):
    """1-core reference matML on the first `sample-sites` columns of OUR data (spawned process so
    the reference's top-level modules never meet the product's).  Returns (dict, lnL of sample)."""
    if not reference_available():
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "unavailable: oracle/_ref not built"}, None
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(1, initializer=_ref_sample_init, initargs=(a.taxa, a.patterns, codes_sample)) as pool:
        pool.map(_ref_eval, [0])
        res = pool.map(_ref_eval, [0, 1, 2], chunksize=1)
    sec = statistics.median(r[0] for r in res)
    n = codes_sample.shape[1]
    return {"value": 1.0 / (sec * a.patterns / n), "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"first {n} of {a.patterns} sites, median of 3 matML calls ({sec:.3f} s each), extrapolated "
                      "linearly in the site count (sites are independent)"}, res[0][1]


def _ref_sample_init(taxa, patterns, codes_sample):
    _ref_worker_init(taxa, patterns, BLOCK, None)
    aln, rconfig, rmcmc = _REF["aln"], _REF["rconfig"], _REF["rmcmc"]
    eye = np.eye(2)
    rconfig.N_TAXA, rconfig.N_CHARS, rconfig.N_SITES = taxa, 2, codes_sample.shape[1]
    rconfig.MODEL, rconfig.IN_DTYPE = "GTR", "bin"
    _REF["leaves"] = {t + 1: np.ascontiguousarray(eye[codes_sample[t]].T) for t in range(taxa)}
    pi, er = aln.pi.copy(), aln.er.copy()
    _REF["tmats"] = [rmcmc.get_prob_t(pi, aln.tree, er, r) for r in aln.rates]
    _REF["edges"] = aln.edge_order()
    _REF["pi"] = pi


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs WORLD_SIZE={a.gpus} (launch through torch.distributed.run)")
    dist = None
    if world > 1:
        import torch.distributed as dist  # host-side rendezvous only (gloo); no torch on the device path
        dist.init_process_group("gloo")

    from cybayes_b200 import _lib, config, likelihood
    from cybayes_b200.alignment import LeafMatrices
    from cybayes_b200.ML_gamma import matML
    from cybayes_b200.subst import gtr_eigensystem
    from cybayes_b200.synthetic import SyntheticAlignment, shard_bounds

    os.environ["CYBAYES_DEVICE"] = str(local)
    aln = SyntheticAlignment(a.taxa, a.patterns, 2, SEED, block_sites=BLOCK)
    lo, hi = shard_bounds(a.patterns, rank, world, BLOCK)
    t_gen = time.perf_counter()
    codes = aln.codes(lo, hi)
    t_gen = time.perf_counter() - t_gen
    n_local = codes.shape[1]
    C, S, N = 4, 2, a.taxa

    config.N_TAXA, config.N_CHARS, config.N_SITES, config.MODEL, config.IN_DTYPE = N, S, n_local, "GTR", "bin"
    leaves = LeafMatrices(codes, S, np.ones((1, S)))
    config.LEAF_LLMAT = leaves
    eng, _ = likelihood.engine_for(leaves, C)        # uploads the tips once (resident like model weights)
    if world > 1:
        box = [eng.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)

    edges = aln.edge_order()
    plan = likelihood._plan_for(edges)
    ekeys = list(aln.tree.keys())
    n_e = len(ekeys)
    block = eng.alloc_slots(n_e * C)
    slots = np.arange(block.base, block.base + n_e * C, dtype=np.int32)
    d = np.array([aln.tree[e] * r for r in aln.rates for e in ekeys])
    eng.queue_build(_lib.CB_MODEL_GTR_EIG, aln.pi, 0.0, gtr_eigensystem(aln.pi, aln.er), slots, d)  # K1: one launch
    slot_of = {(k, e): block.base + k * n_e + i for k in range(C) for i, e in enumerate(ekeys)}
    pslots = np.array([[slot_of[k, e] for k in range(C)] for e in plan.edge_keys], dtype=np.int32)
    pi = aln.pi

    def step():
        lnl, snap = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=True)
        eng.release_snapshot(snap)
        return lnl

    def barrier():
        if dist is not None:
            dist.barrier()

    for _ in range(a.warmup):
        lnl = step()
    barrier()
    eng.sync()
    clocks = ClockSampler(local) if rank == 0 else None
    st0 = eng.stats()
    kernel_ms = []
    eng.mark(0)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        lnl = step()
        kernel_ms.append(eng.last_eval_ms())
    eng.mark(1)
    eng.sync()
    wall = time.perf_counter() - t0
    dev_ms = eng.mark_elapsed_ms()
    st1 = eng.stats()
    barrier()
    clock_rec = clocks.stop() if clocks else None
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    else:
        wall_ms = wall * 1e3
    ms_per_step = dev_ms / a.steps
    value = 1e3 / ms_per_step

    # roofline of the dominant kernel (prune_s2_kernel<4>): algorithmic bytes of THIS rank's shard per
    # evaluation / device time of the evaluation's launches (CUDA events on the launching stream)
    alg_bytes = 16.0 * C * S * n_local * (N - 2) + 1.0 * N * n_local + 8.0 * n_local
    k_ms = statistics.mean(kernel_ms)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh).get("prune_s2_kernel_dram_bytes_per_eval_1M")
            if traffic is not None and (a.patterns != 1000000 or world != 1):
                traffic = None
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "prune_s2_kernel<4>", "algorithmic_bytes_per_eval": alg_bytes,
                "kernel_ms_per_eval": k_ms, "peak_source": peak_src,
                "launches_per_eval": (st1["kernel_launches"] - st0["kernel_launches"]) / a.steps}
    if traffic is not None:
        # the walk never writes carried / folded partials, so it moves fewer bytes than the algorithmic count: the
        # bandwidth it actually draws (ncu DRAM bytes of this kernel / this run's kernel time) against the same peak
        roofline["traffic_GBps"] = traffic / (k_ms * 1e-3) / 1e9
        roofline["traffic_frac"] = roofline["traffic_GBps"] / peak

    # end to end through the reference-facing call: host P matrices (numpy, one dict per category) ->
    # matML -> float.  Timed region holds the H2D of P matrices + op descriptors and the D2H of lnL.
    host_p = eng.download_pmats(slots).reshape(C, n_e, S, S)
    tm_host = [{e: host_p[k, i] for i, e in enumerate(ekeys)} for k in range(C)]
    args = (n_local, N, C)
    for _ in range(2):
        l2, cache = matML(pi, aln.root, leaves, edges, tm_host, *args)
        del cache
    barrier()
    eng.sync()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        l2, cache = matML(pi, aln.root, leaves, edges, tm_host, *args)
        del cache
    e2e_s = (time.perf_counter() - t0) / a.steps
    s1 = eng.stats()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    assert abs(l2 - lnl) <= 1e-12 * abs(lnl), (l2, lnl)
    e2e = {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) / a.steps,
           "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) / a.steps,
           "note": "tips are resident (uploaded once by the first call, like the reference's LEAF_LLMAT); per step: "
                   "host P matrices + op list in, lnL out"}

    # The same call fed the way this package's own drop-in modules feed it: get_prob_t_all builds the tables on the
    # device from (pi, rates, branch lengths), matML takes the device tables -- what the unchanged reference driver
    # does on the compat modules.  Per step: a few KB of scalars + the op list in, lnL out.  Reported next to `e2e`
    # (which keeps the reference-style HOST matrices as its input), not instead of it.
    from cybayes_b200.subst import get_prob_t_all
    for _ in range(2):
        tabs = get_prob_t_all(pi, aln.tree, aln.er, aln.rates)
        l3, cache = matML(pi, aln.root, leaves, edges, tabs, *args)
        del cache, tabs
    barrier()
    eng.sync()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        tabs = get_prob_t_all(pi, aln.tree, aln.er, aln.rates)
        l3, cache = matML(pi, aln.root, leaves, edges, tabs, *args)
        del cache, tabs
    dev_s = (time.perf_counter() - t0) / a.steps
    s1 = eng.stats()
    if dist is not None:
        import torch
        t = torch.tensor([dev_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t[0])
    assert abs(l3 - lnl) <= 1e-9 * abs(lnl), (l3, lnl)
    e2e["device_built_tables"] = {"value": 1.0 / dev_s, "unit": UNIT,
                                  "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) / a.steps,
                                  "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) / a.steps,
                                  "note": "get_prob_t_all (P(t) built on the device, one launch) + matML per step"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n_taxa": N, "n_patterns": a.patterns, "n_states": S, "n_cats": C,
                   "model": "GTR", "patterns_per_gpu": n_local, "parallelism": f"site-sharded x{world}",
                   "l2": "inputs per step (>= 8 GB of partials per GPU) are far larger than the 126 MB L2; no flush"},
        "lnL": lnl, "wall_ms_per_step": wall_ms / a.steps, "roofline": roofline, "e2e": e2e,
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"], "clocks": clock_rec,
        "data_generation_s": t_gen,
    }

    if rank == 0 and world == 1 and not a.no_extras:
        # dirty-path evaluations (cache_matML's job) on the same alignment: random tip -> root paths
        l_full, snap = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=True)
        parents = {c: p for (p, c) in aln.tree}
        rng = np.random.default_rng(5)
        paths = []
        for tip in rng.integers(1, N + 1, size=24):
            path, n = [], int(tip)
            while n != aln.root:
                n = parents[n]
                path.append(n)
            path.sort(key=plan.index.__getitem__)
            nodes = np.array(path, dtype=np.int32)
            ch = np.array([c for n in path for c in plan.kids[n]], dtype=np.int32)
            ps = np.array([[slot_of[k, (n, c)] for k in range(C)] for n in path for c in plan.kids[n]], dtype=np.int32)
            paths.append((nodes, ch, ps))
        if world == 1:
            ms = []
            for nodes, ch, ps in paths:
                l_d, s2 = eng.eval(snap, nodes, ch, ps, pi, want_snapshot=True)
                ms.append(eng.last_eval_ms())
                eng.release_snapshot(s2)
                assert l_d == l_full
            mean_len = statistics.mean(len(p[0]) for p in paths)
            out["dirty_path"] = {"evals_per_sec": 1e3 / statistics.mean(ms[2:]), "mean_path_nodes": mean_len,
                                 "kernel_ms": statistics.mean(ms[2:]), "launches_per_eval": 1,
                                 "check": "each equals the full-pass lnL bit for bit"}
        eng.release_snapshot(snap)
        # likelihood only (no cache kept): what a proposal that is going to be rejected costs
        ms = []
        for _ in range(6):
            l_only, _ = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=False)
            ms.append(eng.last_eval_ms())
        assert l_only == l_full
        out["lnl_only"] = {"evals_per_sec": 1e3 / statistics.mean(ms[2:]), "kernel_ms": statistics.mean(ms[2:]),
                           "note": "same walk, only the partials that must be read back are written"}

    if rank == 0 and world == 1 and not a.no_extras:
        # CPU baseline beside it: the compiled reference, 1 core, on a bounded sample of the same data,
        # and a parity check of the GPU path against it on that sample.
        ns = min(a.sample_sites, n_local)
        sample = np.ascontiguousarray(codes[:, :ns])
        base, ref_lnl = cpu_baseline_sample(a, sample)
        out["cpu_baseline"] = base
        if ref_lnl is not None:
            from cybayes_b200.engine import Engine
            e2 = Engine(sample, S, C, device=local)
            b2 = e2.alloc_slots(n_e * C)
            sl2 = np.arange(b2.base, b2.base + n_e * C, dtype=np.int32)
            e2.queue_build(_lib.CB_MODEL_GTR_EIG, aln.pi, 0.0, gtr_eigensystem(aln.pi, aln.er), sl2, d)
            got, _ = e2.eval(None, plan.nodes, plan.children, pslots - block.base + b2.base, pi, want_snapshot=False)
            e2.close()
            out["cpu_baseline"]["parity"] = {"sample_lnL_reference": ref_lnl, "sample_lnL_gpu": got,
                                             "rel_err": abs(got - ref_lnl) / abs(ref_lnl)}
        # MCMC generations/s through the driver on the reference's README example (config C1)
        try:
            import io
            from cybayes_b200.driver import run_chain
            likelihood.reset_engines()
            os.environ["CYBAYES_COMPRESS_MAX_SITES"] = "250000"
            likelihood.COMPRESS_MAX_SITES = 250000
            res = run_chain(os.path.join(REPO, "tests", "golden", "data", "narrow.phy"), "F81", 3000, 1000, "bin",
                            os.path.join(tempfile.gettempdir(), "bench_narrow"), out=io.StringIO())
            out["mcmc"] = {"gens_per_sec": res["gens_per_sec"], "config": "C1 narrow.phy F81 bin Gamma-4, 3000 "
                           "generations through cybayes_b200.driver (same trace as the reference driver)",
                           "final_lnL": float(res["state"]["logLikehood"])}
        except Exception as exc:  # extras must never sink the headline number
            out["mcmc"] = {"error": repr(exc)}
    elif rank == 0:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                               "sample": "reported at N=1 only"}

    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

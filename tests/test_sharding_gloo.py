"""CPU, world_size 2 over gloo: the site-sharded evaluation (each rank owns a contiguous slice of
patterns and runs the same op list; one scalar all-reduce) gives the single-rank likelihood."""
import os
import subprocess
import sys

import numpy as np

from conftest import REPO

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
repo = sys.argv[1]
sys.path[:0] = [repo, os.path.join(repo, "tests"), os.path.join(repo, "oracle")]
from fake_engine import FakeEngine
from cybayes_b200.synthetic import SyntheticAlignment, shard_bounds
from cybayes_b200.likelihood import _Plan
import pruning_oracle as oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
aln = SyntheticAlignment(24, 4096, 2, 99, block_sites=512)
lo, hi = shard_bounds(aln.n_sites, rank, world, 512)
codes = aln.codes(lo, hi)
eng = FakeEngine(codes, 2, 4)
edges = aln.edge_order()
plan = _Plan(edges)
tm = [oracle.prob_t("GTR", True, aln.pi, aln.tree, aln.er, r) for r in aln.rates]
keys = list(aln.tree)
eng.upload_pmats(np.arange(4 * len(keys)), np.stack([tm[k][e] for k in range(4) for e in keys]))
slot = {(k, e): k * len(keys) + i for k in range(4) for i, e in enumerate(keys)}
ps = np.array([[slot[k, e] for k in range(4)] for e in plan.edge_keys], dtype=np.int32)
part, _ = eng.eval(None, plan.nodes, plan.children, ps, aln.pi, want_snapshot=False)
t = torch.tensor([part], dtype=torch.float64)
dist.all_reduce(t)           # the one collective of the path (NCCL on the GPUs, gloo here)
if rank == 0:
    full = FakeEngine(aln.codes(0, aln.n_sites), 2, 4)
    full.upload_pmats(np.arange(4 * len(keys)), np.stack([tm[k][e] for k in range(4) for e in keys]))
    want, _ = full.eval(None, plan.nodes, plan.children, ps, aln.pi, want_snapshot=False)
    assert abs(float(t[0]) - want) <= 1e-12 * abs(want), (float(t[0]), want)
    assert hi - lo == 2048
    print("SHARDED_OK", float(t[0]), want)
dist.destroy_process_group()
'''


def test_two_rank_site_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), REPO],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, (res.stdout[-1500:], res.stderr[-3000:])


CHAIN_WORKER = r'''
import os, sys, random
import numpy as np
import torch, torch.distributed as dist
repo = sys.argv[1]
sys.path[:0] = [repo, os.path.join(repo, "tests"), os.path.join(repo, "oracle")]
from fake_engine import FakeEngine
from cybayes_b200 import config
from cybayes_b200.fastchain import NativeChain
from cybayes_b200.synthetic import SyntheticAlignment, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
aln = SyntheticAlignment(16, 1024, 2, 11, block_sites=256)


class ShardedEngine(FakeEngine):
    """This rank's slice of the patterns; eval adds the shard sums in rank order (what the fused cross-GPU sum does)."""
    def eval(self, *args, **kw):
        part, snap = super().eval(*args, **kw)
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([part], dtype=torch.float64))
        total = 0.0
        for p in parts:
            total += float(p[0])
        return total, snap


def run(engine, n):
    random.seed(77); np.random.seed(77)
    config.N_TAXA, config.N_CHARS, config.MODEL, config.IN_DTYPE, config.N_CATS = 16, 2, "GTR", "bin", 4
    state = {"tree": dict(aln.tree), "pi": aln.pi.copy(), "rates": aln.er.copy(), "srates": float(aln.alpha), "root": aln.root}
    chain = NativeChain(engine, state, list(aln.rates), "GTR", True, use_callbacks=True, skip_degenerate_rates=True)
    chain.take_rng()
    out = chain.run(n)
    st = chain.state()
    chain.close()
    return out, st

lo, hi = shard_bounds(aln.n_sites, rank, world, 256)
(mv, acc, cur, prop, ratio, logu), st = run(ShardedEngine(aln.codes(lo, hi), 2, 4), 150)
# every rank took the same moves and decisions and ends in the same state ...
digest = torch.tensor([float(mv.sum()), float(acc.sum()), float(cur[-1]), float(st["logLikehood"]), sum(st["tree"].values())],
                      dtype=torch.float64)
every = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(every, digest)
assert all(torch.equal(e, every[0]) for e in every), every
if rank == 0:
    # ... which is the chain of one rank holding the whole alignment (sums of shard sums differ from the sum over all
    # sites in the last bits: decisions must agree, likelihoods to 1e-12)
    (mv1, acc1, cur1, prop1, _, _), st1 = run(FakeEngine(aln.codes(0, aln.n_sites), 2, 4), 150)
    assert np.array_equal(mv, mv1) and np.array_equal(acc, acc1) and acc.sum() > 5
    assert np.allclose(cur, cur1, rtol=1e-12, atol=0) and np.allclose(prop, prop1, rtol=1e-12, atol=0)
    assert list(st["tree"]) == list(st1["tree"])
    print("SHARDED_CHAIN_OK", int(acc.sum()), float(st["logLikehood"]))
dist.destroy_process_group()
'''


def test_two_rank_sharded_native_chain(tmp_path):
    """The generation loop of the library on two ranks, each holding half of the patterns (bench.py's
    mcmc_on_workload under torchrun): same moves, decisions and final state on both ranks and as on one rank."""
    script = tmp_path / "chain_worker.py"
    script.write_text(CHAIN_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script), REPO],
                         capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0 and "SHARDED_CHAIN_OK" in res.stdout, (res.stdout[-1500:], res.stderr[-3000:])


def test_shard_bounds_cover_everything():
    from cybayes_b200.synthetic import shard_bounds
    for n, g in ((1000000, 125000), (1000, 64), (5, 64), (129, 64)):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(n, r, world, g) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(lo % g == 0 or lo == n for lo, _ in cuts)

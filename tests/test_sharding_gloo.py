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


def test_shard_bounds_cover_everything():
    from cybayes_b200.synthetic import shard_bounds
    for n, g in ((1000000, 125000), (1000, 64), (5, 64), (129, 64)):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(n, r, world, g) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(lo % g == 0 or lo == n for lo, _ in cuts)

"""GPU, >= 2 devices (skipped on a one-GPU box): the restated driver with the alignment sharded over two
ranks (CYBAYES_SHARD=1, one process per GPU, scalar NCCL all-reduce inside the library) reproduces the
reference's recorded accept/reject trace on every rank."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO, load_trace

pytestmark = pytest.mark.gpu

WORKER = r'''
import io, json, os, sys
repo, out_dir = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
os.environ["CYBAYES_DEVICE"] = os.environ["LOCAL_RANK"]
from cybayes_b200.driver import run_chain
rank = int(os.environ["RANK"])
rec = []
res = run_chain(os.path.join(repo, "tests", "golden", "data", "narrow.phy"), "F81", 1000, 1, "bin",
                os.path.join(out_dir, f"run{rank}"), out=io.StringIO(),
                on_generation=lambda i, cur, prop, p, mv, acc, st: rec.append([i, float(cur), float(prop), str(p), mv]))
json.dump({"init": float(res["initial_lnL"]), "rec": rec}, open(os.path.join(out_dir, f"trace{rank}.json"), "w"))
'''


def _device_count():
    import ctypes
    from cybayes_b200 import _lib
    n = ctypes.c_int(0)
    return n.value if _lib.load().cb_device_count(ctypes.byref(n)) == 0 else 0


def test_sharded_driver_trace(tmp_path):
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CYBAYES_SHARD="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29633")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", str(script), REPO, str(tmp_path)],
                         capture_output=True, text=True, env=env, timeout=280)
    assert res.returncode == 0, res.stderr[-3000:]
    rows, meta = load_trace("narrow_F81")
    for rank in (0, 1):
        got = json.load(open(tmp_path / f"trace{rank}.json"))
        assert abs(got["init"] - meta["init_lnL"]) <= 1e-11 * abs(meta["init_lnL"])
        for r, g in zip(got["rec"], rows):
            assert (r[3], r[4]) == (g["param"], g["move"]), (rank, r, g)
            assert abs(r[2] - float(g["proposed_ll"])) <= 1e-11 * abs(float(g["proposed_ll"])), (rank, r, g)

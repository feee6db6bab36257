"""CPU: the native generation loop (csrc/mcmc_native.cuh through cybayes_b200.fastchain) with the likelihood side
replaced by callbacks into the oracle-backed fake engine.  What is checked is everything the loop itself owns: both
random streams, the tree-dict order semantics, the proposal arithmetic, the op lists it builds, the bookkeeping of P
slots and snapshots -- against the traces recorded from the unmodified reference driver."""
import io
import random

import numpy as np
import pytest

import golden_io
from conftest import load_trace
from trace_checks import check_outputs, compare_trace

CPU_TRACES = [("binary_F81", 300), ("twoStates_JC", 300), ("narrow_F81", 600), ("narrow_JC", 300), ("phon_ringe_F81", 300),
              ("phon_ringe_GTR", 150), ("ie42_JC", 60)]


@pytest.mark.parametrize("name,n_gen", CPU_TRACES)
def test_native_chain_reproduces_reference_trace(name, n_gen, golden_cases, fake_backend, tmp_path):
    from cybayes_b200.fastchain import run_chain_native
    case = golden_cases[name]
    rows, meta = load_trace(name)
    rec = []
    res = run_chain_native(golden_io.data_path(case), case["model"], n_gen, 1, case["dtype"], str(tmp_path / "run"),
                           out=io.StringIO(),
                           on_generation=lambda i, cur, prop, p, mv, acc, st: rec.append((i, cur, prop, p, mv, acc)))
    compare_trace(rec, rows[:n_gen], meta, res["initial_lnL"], 1e-9 if case["model"] == "GTR" else 1e-11)
    if n_gen == len(rows):
        check_outputs(str(tmp_path / "run"), rows, meta, res)
    else:   # tree length and alpha of every generation: exact strings
        for lr, g in zip(open(str(tmp_path / "run.log")).read().splitlines()[1:], rows):
            f = lr.split("\t")
            assert f[2] == g["log_TL"] and f[3] == g["alpha"], (f, g)
    # accepted flags are consistent with the recorded chain: the state lnL after generation i is the next current lnL
    for (i, cur, prop, p, mv, acc), nxt in zip(rec, rec[1:]):
        assert nxt[1] == (prop if acc else cur), i


def test_native_chain_hands_the_generators_back(golden_cases, fake_backend, tmp_path):
    """After a native run the interpreter's two generators continue exactly where the Python driver's would."""
    from cybayes_b200 import likelihood
    from cybayes_b200.driver import run_chain
    from cybayes_b200.fastchain import run_chain_native
    case = golden_cases["phon_ringe_F81"]
    run_chain_native(golden_io.data_path(case), "F81", 137, 50, "multi", str(tmp_path / "a"), out=io.StringIO())
    got = (random.random(), random.getrandbits(32), float(np.random.random_sample()))
    likelihood.reset_engines()
    run_chain(golden_io.data_path(case), "F81", 137, 50, "multi", str(tmp_path / "b"), out=io.StringIO(), fast_spr=True)
    want = (random.random(), random.getrandbits(32), float(np.random.random_sample()))
    assert got == want
    assert open(str(tmp_path / "a.log")).read().split("\t")[2::3] == open(str(tmp_path / "b.log")).read().split("\t")[2::3]
    assert open(str(tmp_path / "a.trees")).read() == open(str(tmp_path / "b.trees")).read()


def test_native_chain_reports_the_reference_crash(golden_cases, fake_backend, tmp_path):
    """GTR on binary data has one exchangeability: the reference dies in random.sample(range(1), 2) at the first
    `rates` move; the native loop reports the same condition instead of sampling out of range."""
    from cybayes_b200._lib import CyBayesB200Error
    from cybayes_b200.fastchain import run_chain_native
    case = golden_cases["narrow_GTR"]
    with pytest.raises(CyBayesB200Error, match="Sample larger than population"):
        run_chain_native(golden_io.data_path(case), "GTR", 400, 100, "bin", str(tmp_path / "g"), out=io.StringIO())


def test_gtr_on_binary_data_runs_with_the_degenerate_block_skipped(golden_cases, fake_backend, tmp_path):
    """with skip_degenerate_rates the single exchangeability of a 2-state GTR model is never
    proposed, so the chain runs -- the Python driver and the native loop agree generation by generation."""
    from cybayes_b200 import likelihood
    from cybayes_b200.driver import run_chain
    from cybayes_b200.fastchain import run_chain_native
    case = golden_cases["narrow_GTR"]
    a, b = [], []
    run_chain(golden_io.data_path(case), "GTR", 120, 40, "bin", str(tmp_path / "p"), out=io.StringIO(), fast_spr=True,
              skip_degenerate_rates=True, on_generation=lambda i, cur, prop, p, mv, acc, st: a.append((p, mv, acc, float(prop))))
    likelihood.reset_engines()
    run_chain_native(golden_io.data_path(case), "GTR", 120, 40, "bin", str(tmp_path / "n"), out=io.StringIO(),
                     skip_degenerate_rates=True, on_generation=lambda i, cur, prop, p, mv, acc, st: b.append((p, mv, acc, float(prop))))
    assert len(a) == 120 and not any(p == "rates" for p, _, _, _ in a)
    assert [x[:3] for x in a] == [x[:3] for x in b]
    np.testing.assert_allclose([x[3] for x in a], [x[3] for x in b], rtol=1e-12)
    assert open(str(tmp_path / "p.trees")).read() == open(str(tmp_path / "n.trees")).read()


def test_likelihood_first_policy_keeps_the_trace(golden_cases, fake_backend, tmp_path, monkeypatch):
    """Full-table proposals (pi / alpha) evaluated without a cache and repeated with one on acceptance
    (CYBAYES_CHAIN_LNL_FIRST=1; automatic on large alignments once such proposals are rarely accepted): same trace
    as the reference, one extra evaluation per accepted full-table proposal and no snapshot left behind."""
    from cybayes_b200.fastchain import run_chain_native
    case = golden_cases["narrow_F81"]
    rows, meta = load_trace("narrow_F81")
    n_gen = 400

    def run(tag):
        rec = []
        res = run_chain_native(golden_io.data_path(case), "F81", n_gen, 1, "bin", str(tmp_path / tag), out=io.StringIO(),
                               on_generation=lambda i, cur, prop, p, mv, acc, st: rec.append((i, cur, prop, p, mv, acc)))
        eng = fake_backend.instances[-1]
        return rec, res, eng.n_evals, len(eng.snaps)

    rec0, res0, evals0, snaps0 = run("plain")
    monkeypatch.setenv("CYBAYES_CHAIN_LNL_FIRST", "1")
    from cybayes_b200 import likelihood
    likelihood.reset_engines()
    rec1, res1, evals1, snaps1 = run("first")
    compare_trace(rec1, rows[:n_gen], meta, res1["initial_lnL"], 1e-11)
    assert rec1 == rec0
    accepted_full = sum(1 for (_, _, _, p, _, acc) in rec1 if acc and p in ("pi", "srates"))
    assert accepted_full > 0 and evals1 == evals0 + accepted_full
    assert snaps1 == snaps0
    assert open(str(tmp_path / "plain.trees")).read() == open(str(tmp_path / "first.trees")).read()

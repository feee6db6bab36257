"""TEST INFRASTRUCTURE: comparisons of a chain run against the per-generation traces recorded from the unmodified
reference driver (tests/golden/traces/*.tsv, tests/golden/make_golden.py)."""
import numpy as np

TRACES = [("binary_F81", 300), ("twoStates_F81", 300), ("twoStates_JC", 300), ("narrow_F81", 3000),
          ("narrow_JC", 500), ("broad_F81", 300), ("phon_ringe_JC", 500), ("phon_ringe_F81", 500),
          ("phon_ringe_GTR", 300), ("ie42_JC", 200), ("ie42_GTR", 60)]


def compare_trace(rec, rows, meta, init_lnl, rel):
    """rec: (iter, current lnL, proposed lnL, param, move, accepted) per generation."""
    assert abs(init_lnl - meta["init_lnL"]) <= rel * abs(meta["init_lnL"])
    assert len(rec) == len(rows)
    for r, g in zip(rec, rows):
        i, cur, prop, param, move = r[:5]
        margin = f"first divergent generation {i}: got {r}, reference {g}"
        assert str(param) == g["param"] and move == g["move"], margin
        want = float(g["proposed_ll"])
        if np.isfinite(want):
            assert abs(prop - want) <= rel * abs(want), margin
        assert abs(cur - float(g["current_ll"])) <= rel * abs(float(g["current_ll"])), margin


def check_outputs(prefix, rows, meta, res):
    """.log (tree length and alpha: exact strings), .trees (final sampled tree: exact Newick string), counters."""
    log_rows = open(prefix + ".log").read().splitlines()[1:]
    assert len(log_rows) == len(rows)
    for lr, g in zip(log_rows, rows):
        f = lr.split("\t")
        assert f[2] == g["log_TL"] and f[3] == g["alpha"], (f, g)
    trees = open(prefix + ".trees").read().strip().splitlines()
    assert trees[-1].split("\t")[1] == meta["last_tree"]
    counters = sorted(f"({str(k[0])!r}, {k[1]!r}) {res['accepts'].get(k, 0)} {v}" for k, v in res["moves"].items())
    assert counters == sorted(c.replace("np.str_(", "").replace("'),", "',", 1) for c in meta["counters"])

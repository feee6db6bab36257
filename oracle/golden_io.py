"""TEST INFRASTRUCTURE: load the committed golden cases (tests/golden/cases.json) and rebuild, for
one case, the inputs both sides need: the alignment through the oracle's reader restatement, the
start tree / pi / rates / category rates exactly as recorded from the unmodified reference."""
import json
import os

import numpy as np

import pruning_oracle as oracle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def load_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as fh:
        return {c["name"]: c for c in json.load(fh)["cases"]}


def data_path(case):
    return os.path.join(GOLDEN, "data", case["file"])


def case_state(case):
    """(tree dict in recorded insertion order, pi, rates or None, edge order, site rates)."""
    tree = {(p, c): t for p, c, t in case["tree"]}
    edges = [tuple(e) for e in case["postorder"]]
    rates = None if case["rates"] is None else np.array(case["rates"])
    return tree, np.array(case["pi"]), rates, edges, list(case["site_rates"])


def oracle_lnl(case, tree=None, gtr_via="expm", scaled=False):
    """lnL of the case's start state (or of `tree`) computed by the NumPy oracle end to end."""
    tree0, pi, rates, edges, site_rates = case_state(case)
    tree = tree0 if tree is None else tree
    _, S, _, _, ll, _, n_sites = oracle.read_phylip(data_path(case), case["reader"])
    tm = [oracle.prob_t(case["model"], case["dtype"] == "bin", pi, tree, rates, r, beta=case["norm_beta"],
                        gtr_via=gtr_via) for r in site_rates]
    if scaled:
        return oracle.mat_ml_scaled(pi, case["root"], ll, edges, tm, n_sites, case["n_taxa"])
    return oracle.mat_ml(pi, case["root"], ll, edges, tm, n_sites, case["n_taxa"])[0]

"""CPU: the NumPy oracle against the golden vectors recorded from the unmodified reference, and
against the compiled reference itself (oracle/_ref) when it is present."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_io
import pruning_oracle as oracle
from conftest import REPO

FAST = ["binary_F81", "twoStates_F81", "twoStates_JC", "narrow_F81", "narrow_GTR", "phon_ringe_JC", "phon_ringe_F81",
        "phon_ringe_GTR", "ie42_JC", "ielex_multistate_F81"]


def _leaf_digest(ll, n_taxa):
    h = hashlib.sha256()
    for k in range(1, n_taxa + 1):
        h.update(np.packbits(np.ascontiguousarray(ll[k]).astype(bool), axis=None).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_vectors(name, golden_cases):
    case = golden_cases[name]
    n_leaves, S, alphabet, _, ll, taxa, n_sites = oracle.read_phylip(golden_io.data_path(case), case["reader"])
    # integer artefacts: exact
    assert (n_leaves, S, n_sites) == (case["n_taxa"], case["n_chars"], case["n_sites"])
    assert alphabet == case["alphabet"] and taxa == case["taxa"]
    assert _leaf_digest(ll, n_leaves) == case["leaf_digest"]
    tree, pi, rates, edges, site_rates = golden_io.case_state(case)
    assert oracle.edge_order(tree, case["root"], n_leaves) == edges
    np.testing.assert_array_equal(oracle.site_rates(case["srates"]), site_rates)
    if case["model"] == "F81":
        assert oracle.f81_beta(pi) == case["norm_beta"]
    # P(t) and likelihood
    tm = [oracle.prob_t(case["model"], case["dtype"] == "bin", pi, tree, rates, r, beta=case["norm_beta"])
          for r in site_rates]
    for key, mats in case["pmat_sample"].items():
        e = tuple(int(x) for x in key.split(","))
        for k, ref in enumerate(mats):
            if isinstance(ref, dict):
                assert np.array_equal(tm[k][e][0], ref["row0"]) and np.array_equal(np.diag(tm[k][e]), ref["diag"])
            else:
                assert np.array_equal(tm[k][e], np.array(ref))
    lnl, cache = oracle.mat_ml(pi, case["root"], ll, edges, tm, n_sites, n_leaves)
    assert abs(lnl - case["lnL"]) <= 1e-14 * abs(case["lnL"])
    for node, sums in case["partial_sums"].items():
        np.testing.assert_allclose([cache[k][int(node)].sum() for k in range(4)], sums, rtol=1e-13)
    scaled = oracle.mat_ml_scaled(pi, case["root"], ll, edges, tm, n_sites, n_leaves)
    assert abs(scaled - case["lnL"]) <= 1e-13 * abs(case["lnL"])
    parents = oracle.parent_of(tree)
    for d in case["dirty"]:
        e = tuple(d["edge"])
        assert oracle.path_to_root(parents, e[1], case["root"]) == d["path"]
        saved = [tm[k][e] for k in range(4)]
        Q = oracle.gtr_q(rates, pi) if case["model"] == "GTR" else None
        for k, r in enumerate(site_rates):
            tm[k][e] = oracle.p_matrix(case["model"], case["dtype"] == "bin", pi, rates,
                                       oracle.f81_beta(pi) if case["model"] == "F81" else case["norm_beta"],
                                       d["new_t"] * r, Q=Q)
        l2, _ = oracle.cache_mat_ml(pi, case["root"], ll, cache, d["path"], edges, tm, n_sites, n_leaves)
        assert abs(l2 - d["lnL"]) <= 1e-14 * abs(d["lnL"])
        for k in range(4):
            tm[k][e] = saved[k]


def test_scaled_oracle_survives_where_the_reference_underflows():
    rng = np.random.default_rng(3)
    n_taxa, n_sites = 1300, 20
    tree, prev, nxt = {}, 1, n_taxa + 1
    for tip in range(2, n_taxa + 1):
        tree[nxt, prev] = 0.6
        tree[nxt, tip] = 0.6
        prev, nxt = nxt, nxt + 1
    root = nxt - 1
    kids = oracle.children_of(tree)
    order, stack = [], [root]
    while stack:
        nd = stack.pop()
        x, y = kids[nd]
        order += [(nd, x), (nd, y)]
        stack += [c for c in (y, x) if c > n_taxa]
    edges = order[::-1]
    pi = np.array([0.4, 0.6])
    eye = np.eye(2)
    ll = {t: np.ascontiguousarray(eye[rng.integers(0, 2, n_sites)].T) for t in range(1, n_taxa + 1)}
    tm = [oracle.prob_t("F81", True, pi, tree, None, r) for r in oracle.site_rates(0.5)]
    assert oracle.mat_ml(pi, root, ll, edges, tm, n_sites, n_taxa)[0] == -np.inf
    total, per_site = oracle.mat_ml_scaled(pi, root, ll, edges, tm, n_sites, n_taxa, site_lnl=True)
    assert np.isfinite(total) and (per_site < -709).all()


def test_gtr_eigen_route_matches_expm():
    rng = np.random.default_rng(11)
    for S in (2, 6, 23):
        pi = rng.dirichlet(np.full(S, 3.0))
        er = rng.dirichlet(np.ones(S * (S - 1) // 2))
        for d in (1e-4, 0.02, 0.5, 3.0):
            a = oracle.p_matrix("GTR", False, pi, er, None, d, gtr_via="expm")
            b = oracle.p_matrix("GTR", False, pi, er, None, d, gtr_via="eig")
            np.testing.assert_allclose(b, a, rtol=5e-10, atol=1e-15)


@pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "oracle", "_ref")), reason="oracle/_ref not built")
def test_oracle_against_compiled_reference_on_random_inputs():
    """Fresh interpreter (the reference's top-level module names must not leak into this one):
    random tree / data, the compiled reference's get_prob_t + matML + cache_matML vs the oracle."""
    code = r'''
import sys, random
import numpy as np
sys.path.insert(0, "oracle/_ref"); sys.path.insert(0, "oracle")
import config, mcmc_gamma, ML_gamma
import pruning_oracle as oracle
rng = np.random.default_rng(5)
for S, model, dtype in ((2, "F81", "bin"), (5, "F81", "multi"), (5, "JC", "multi"), (4, "GTR", "multi")):
    n_taxa, n_sites = 11, 57
    config.N_TAXA, config.N_CHARS, config.N_SITES, config.MODEL, config.IN_DTYPE = n_taxa, S, n_sites, model, dtype
    nodes, nxt, tree = list(range(1, n_taxa + 1)), n_taxa + 1, {}
    while len(nodes) > 1:
        a = nodes.pop(int(rng.integers(len(nodes)))); b = nodes.pop(int(rng.integers(len(nodes))))
        tree[nxt, a] = float(rng.exponential(0.1)); tree[nxt, b] = float(rng.exponential(0.1)); nodes.append(nxt); nxt += 1
    root = nxt - 1
    pi = rng.dirichlet(np.ones(S)) if model != "JC" else np.repeat(1.0 / S, S)
    er = rng.dirichlet(np.ones(S * (S - 1) // 2))
    config.NORM_BETA = 1 / (1 - np.dot(pi, pi))
    eye = np.vstack([np.eye(S), np.ones(S)])
    ll = {t: np.ascontiguousarray(eye[rng.integers(0, S + 1, n_sites)].T) for t in range(1, n_taxa + 1)}
    edges = mcmc_gamma.postorder(mcmc_gamma.adjlist2nodes_dict(tree), root)[::-1]
    assert edges == oracle.edge_order(tree, root, n_taxa)
    rates = mcmc_gamma.get_siterates(0.8)
    assert rates == oracle.site_rates(0.8)
    tm_ref = [mcmc_gamma.get_prob_t(pi, tree, er, r) for r in rates]
    tm = [oracle.prob_t(model, dtype == "bin", pi, tree, er, r, beta=config.NORM_BETA) for r in rates]
    for k in range(4):
        for e in tree:
            assert np.array_equal(np.asarray(tm_ref[k][e]), tm[k][e]), (model, e)
    l_ref, c_ref = ML_gamma.matML(pi, root, ll, edges, tm_ref, n_sites, n_taxa, 4)
    l_or, c_or = oracle.mat_ml(pi, root, ll, edges, tm, n_sites, n_taxa)
    assert l_ref == l_or, (l_ref, l_or)
    path = mcmc_gamma.get_path2root(mcmc_gamma.adjlist2reverse_nodes_dict(tree), 3, root)
    assert path == oracle.path_to_root(oracle.parent_of(tree), 3, root)
    l2_ref, _ = ML_gamma.cache_matML(pi, root, ll, c_ref, path, edges, tm_ref, n_sites, n_taxa, 4)
    l2_or, _ = oracle.cache_mat_ml(pi, root, ll, c_or, path, edges, tm, n_sites, n_taxa)
    assert l2_ref == l2_or == l_ref
print("OK")
'''
    res = subprocess.run([sys.executable, "-c", code], cwd=REPO, capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr[-2000:]

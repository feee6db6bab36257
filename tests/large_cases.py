"""TEST INFRASTRUCTURE: bodies of tests/test_gpu_parity_large.py (GPU) and of their small CPU dry runs on
the oracle-backed fake engine (tests/test_host_logic.py).

GPU parity at the benchmark shapes (-m gpu): the kernel instantiations that produce the headline numbers,
checked against the NumPy oracle (oracle/pruning_oracle.py: mat_ml_scaled, ML_gamma.pyx:7-42 + rescaling).

* C4 shape -- 1024 taxa x 131 072 simulated binary sites, GTR + Gamma-4: from 75 776 patterns the 2-state
  family runs prune_s2_kernel<4, 1, 256, 3>; a single-launch depth-first walk (what one GPU runs at 1M
  patterns), the two-launch split walk (what 4- and 8-GPU shards run, and the default at this size), the
  likelihood-only variants, a dirty path and cached partials (stored and folded-cherry nodes).
* C5 shape -- 256 taxa x 64 states x 8 192 simulated sites, GTR + Gamma-4: the register-carried FP64 tensor
  kernel prune_dmma_rc_kernel<64, true> in its walk schedules.

Tolerance: both sides get the SAME transition matrices (scipy expm, as the reference builds them,
mcmc_gamma.pyx:481), so lnL must agree to <= 1e-11 relative (summation order only) and partials to 1e-12;
the device-built GTR matrices (eigendecomposition) are held to BASELINE.json's 1e-9.
"""
import numpy as np

import pruning_oracle as oracle

SEED = 20260101
REL = 1e-11


def _setup(n_taxa, n_sites, n_states, seed, block):
    from cybayes_b200 import config
    from cybayes_b200.alignment import LeafMatrices
    from cybayes_b200.synthetic import SyntheticAlignment
    aln = SyntheticAlignment(n_taxa, n_sites, n_states, seed, block_sites=block)
    codes = aln.codes(0, n_sites)
    config.N_TAXA, config.N_CHARS, config.N_SITES = n_taxa, n_states, n_sites
    config.MODEL, config.IN_DTYPE, config.N_CATS = "GTR", ("bin" if n_states == 2 else "multi"), 4
    leaves = LeafMatrices(codes, n_states, np.ones((1, n_states)))
    config.LEAF_LLMAT = leaves
    return aln, codes, leaves


def _engine_inputs(eng, plan, tree, tm, C):
    ekeys = list(tree.keys())
    n_e = len(ekeys)
    block = eng.alloc_slots(n_e * C)
    slot_of = {(k, e): block.base + k * n_e + i for k in range(C) for i, e in enumerate(ekeys)}
    eng.upload_pmats(np.arange(block.base, block.base + n_e * C, dtype=np.int32),
                     np.stack([tm[k][e] for k in range(C) for e in ekeys]))
    pslots = np.array([[slot_of[k, e] for k in range(C)] for e in plan.edge_keys], dtype=np.int32)
    return block, slot_of, pslots


def _check_partial(eng, snap, node, kept):
    got_m, got_e = eng.read_partial(snap, node, with_scale=True)
    want_m, want_e = kept[node]
    # same value m * 2^e; the general family keeps one exponent per category, so compare on the oracle's exponent
    np.testing.assert_allclose(np.ldexp(got_m, (got_e - want_e)[None, None, :]), want_m, rtol=1e-12, atol=0)


def run_c4_shape(N, P, chunk):
    """prune_s2_kernel<4,1,256,3>: single-launch walk, split walk, likelihood-only, dirty path, cached partials."""
    from cybayes_b200.likelihood import _plan_for, engine_for
    from cybayes_b200.ML_gamma import cache_matML, matML
    from cybayes_b200.mcmc_gamma import get_prob_t
    S, C = 2, 4
    aln, codes, leaves = _setup(N, P, S, SEED, chunk)
    edges = aln.edge_order()
    pi, root = aln.pi, aln.root
    tm = [oracle.prob_t("GTR", True, pi, aln.tree, aln.er, r) for r in aln.rates]      # scipy expm, like the reference
    plan = _plan_for(edges)
    kids = plan.kids
    heavy = max(kids[root])                      # an internal child of the root (ids grow towards the root)
    cherry = next(n for n in plan.node_list if all(c <= N for c in kids[n]))   # a folded cherry: materialised on read
    mid = plan.node_list[len(plan.node_list) // 2]
    keep_nodes = tuple(n for n in {heavy, cherry, mid} if n > N)
    want, _, kept = oracle.mat_ml_scaled_codes(pi, root, codes, S, None, edges, tm, N, chunk=chunk,
                                               keep_nodes=keep_nodes)
    assert np.isfinite(want)

    # (1) the reference-facing call with reference-style HOST matrices (default schedule at this size: split walk)
    lnl, cache = matML(pi, root, leaves, edges, tm, P, N, C)
    assert abs(lnl - want) <= REL * abs(want), (lnl, want)
    eng, _ = engine_for(leaves, C)
    assert eng.n_patterns == P                   # no compression: the kernel really sees 131 072 patterns
    for node in keep_nodes:
        _check_partial(eng, cache.snap, node, kept)

    # (2) every schedule of the big instantiation, engine level: bit-identical to each other, 1e-11 to the oracle
    block, slot_of, pslots = _engine_inputs(eng, plan, aln.tree, tm, C)
    l_walk, s_walk = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=True, force_walk=True)
    l_walk_only, _ = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=False, force_walk=True)
    l_split, s_split = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=True)
    l_split_only, _ = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=False)
    assert abs(l_walk - want) <= REL * abs(want), (l_walk, want)
    assert l_walk == l_walk_only == l_split == l_split_only == float(lnl)
    for node in keep_nodes:
        _check_partial(eng, s_walk, node, kept)
        _check_partial(eng, s_split, node, kept)
    eng.release_snapshot(s_split)

    # (3) dirty path (cache_matML, ML_gamma.pyx:83-118): one tip edge takes another edge's matrices
    tip = 7
    parents = oracle.parent_of(aln.tree)
    path = oracle.path_to_root(parents, tip, root)
    e_tip, e_other = (parents[tip], tip), list(aln.tree.keys())[5]
    tm2 = [dict(t) for t in tm]
    for k in range(C):
        tm2[k][e_tip] = tm[k][e_other]
    want2, _, _ = oracle.mat_ml_scaled_codes(pi, root, codes, S, None, edges, tm2, N, chunk=chunk)
    l2, cache2 = cache_matML(pi, root, leaves, cache, path, edges, tm2, P, N, C)
    assert abs(l2 - want2) <= REL * abs(want2), (l2, want2)
    assert l2 != lnl
    # ... and the same dirty path from the single-launch walk's snapshot, engine level: same bits
    todo = sorted(path, key=plan.index.__getitem__)
    nodes = np.array(todo, dtype=np.int32)
    ch = np.array([c for n in todo for c in kids[n]], dtype=np.int32)
    ps = np.array([[slot_of[k, e_other if (n, c) == e_tip else (n, c)] for k in range(C)]
                   for n in todo for c in kids[n]], dtype=np.int32)
    l3, _ = eng.eval(s_walk, nodes, ch, ps, pi, want_snapshot=False)
    assert l3 == float(l2)
    eng.release_snapshot(s_walk)

    # (4) device-built GTR matrices (cb_pmat_build, eigendecomposition): BASELINE.json's 1e-9
    tabs = [get_prob_t(pi, aln.tree, aln.er, r) for r in aln.rates]
    l4, _ = matML(pi, root, leaves, edges, tabs, P, N, C)
    assert abs(l4 - want) <= 1e-9 * abs(want), (l4, want)




def run_c5_shape(N, P, S, chunk):
    """prune_dmma_rc_kernel<64, true> on a 256-taxon x 64-state x 8192-site simulated alignment."""
    from cybayes_b200.likelihood import _plan_for, engine_for
    from cybayes_b200.ML_gamma import cache_matML, matML
    C = 4
    aln, codes, leaves = _setup(N, P, S, 20260102, chunk)
    rng = np.random.default_rng(3)
    codes = codes.copy()
    codes[rng.random(codes.shape) < 0.05] = S          # missing cells ('?'): the all-ones ambiguity set
    leaves.codes = codes
    edges = aln.edge_order()
    pi, root = aln.pi, aln.root
    tm = [oracle.prob_t("GTR", False, pi, aln.tree, aln.er, r) for r in aln.rates]
    plan = _plan_for(edges)
    kids = plan.kids
    heavy = max(kids[root])
    mid = plan.node_list[len(plan.node_list) // 2]
    keep_nodes = tuple({heavy, mid})
    want, _, kept = oracle.mat_ml_scaled_codes(pi, root, codes, S, None, edges, tm, N, chunk=chunk, keep_nodes=keep_nodes)
    assert np.isfinite(want)
    lnl, cache = matML(pi, root, leaves, edges, tm, P, N, C)
    assert abs(lnl - want) <= REL * abs(want), (lnl, want)
    eng, _ = engine_for(leaves, C)
    assert eng.n_patterns == P
    for node in keep_nodes:
        _check_partial(eng, cache.snap, node, kept)
    block, slot_of, pslots = _engine_inputs(eng, plan, aln.tree, tm, C)
    l_walk, s_walk = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=True, force_walk=True)
    l_only, _ = eng.eval(None, plan.nodes, plan.children, pslots, pi, want_snapshot=False)
    assert l_walk == l_only == float(lnl)
    for node in keep_nodes:
        _check_partial(eng, s_walk, node, kept)
    eng.release_snapshot(s_walk)
    # dirty path
    tip = 3
    parents = oracle.parent_of(aln.tree)
    path = oracle.path_to_root(parents, tip, root)
    e_tip, e_other = (parents[tip], tip), list(aln.tree.keys())[9]
    tm2 = [dict(t) for t in tm]
    for k in range(C):
        tm2[k][e_tip] = tm[k][e_other]
    want2, _, _ = oracle.mat_ml_scaled_codes(pi, root, codes, S, None, edges, tm2, N, chunk=chunk)
    l2, _ = cache_matML(pi, root, leaves, cache, path, edges, tm2, P, N, C)
    assert abs(l2 - want2) <= REL * abs(want2), (l2, want2)

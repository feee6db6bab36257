"""Synthetic alignments for the benchmark configurations.

Sites are simulated down a random tree with their discrete-Gamma category rate, so the per-site
likelihood at the generating tree stays far above the fp64 underflow that iid-uniform columns hit
in the reference.  Generation is blocked: block b of `block_sites` columns has its own
seed, so any rank can produce exactly its slice of the same alignment whatever the GPU count.
"""
from __future__ import annotations

import numpy as np

from .subst import fnGTR, get_siterates  # noqa: F401  (get_siterates needs config.N_CATS)


def random_tree(n_taxa, rng, bl_mean=0.02):
    """Random-join topology (the scheme of rtree, mcmc_gamma.pyx:305-317, with a NumPy generator),
    node ids as in the reference: tips 1..N, internal N+1.., root = 2N-1.  Returns (tree dict with
    Exp(bl_mean) branch lengths in insertion order, root)."""
    pool = list(range(1, n_taxa + 1))
    rng.shuffle(pool)
    pool = [int(x) for x in pool]
    nxt = n_taxa + 1
    tree = {}
    while len(pool) > 1:
        i = int(rng.integers(len(pool)))
        a = pool.pop(i)
        j = int(rng.integers(len(pool)))
        b = pool.pop(j)
        tree[nxt, a] = float(rng.exponential(bl_mean))
        tree[nxt, b] = float(rng.exponential(bl_mean))
        pool.append(nxt)
        nxt += 1
    return tree, nxt - 1


def transition_matrices(Q, tree, rates):
    """{edge: [P(t r_k) for k]} on the host (simulation only), via the symmetric eigensystem."""
    w, V = np.linalg.eig(Q)
    Vi = np.linalg.inv(V)
    out = {}
    for e, t in tree.items():
        out[e] = [np.clip(np.real((V * np.exp(w * t * r)) @ Vi), 0.0, 1.0) for r in rates]
    return out


def simulate_block(tree, root, n_taxa, pi, pmats, n_cats, n_sites, seed):
    """(n_taxa, n_sites) uint8 states of one block of columns."""
    rng = np.random.default_rng(seed)
    S = len(pi)
    kids = {}
    for (p, c) in tree:
        kids.setdefault(p, []).append(c)
    cat = rng.integers(0, n_cats, size=n_sites)
    out = np.empty((n_taxa, n_sites), dtype=np.uint8)
    root_state = rng.choice(S, size=n_sites, p=pi).astype(np.uint8)
    stack = [(root, root_state)]
    flat = cat * S  # index of (category, parent state) rows
    while stack:
        node, st = stack.pop()
        for child in kids[node]:
            pm = np.stack(pmats[node, child])                             # (C, S, S)
            u = rng.random(n_sites, dtype=np.float32)
            if S == 2:
                flip = np.array([[m[0, 1], m[1, 0]] for m in pm], dtype=np.float32).ravel()
                cs = st ^ (u < flip[flat + st]).astype(np.uint8)
            else:
                cdf = np.cumsum(pm, axis=2).reshape(-1, S).astype(np.float32)
                thresholds = cdf[flat + st]                               # (n_sites, S)
                cs = (u[:, None] > thresholds[:, :-1]).sum(axis=1).astype(np.uint8)
            if child <= n_taxa:
                out[child - 1] = cs
            else:
                stack.append((child, cs))
    return out


class SyntheticAlignment:
    """Deterministic description of a benchmark alignment; `codes(lo, hi)` materialises columns
    [lo, hi) (block-aligned) on the host."""

    def __init__(self, n_taxa, n_sites, n_states, seed, alpha=0.5, bl_mean=0.02, n_cats=4, block_sites=125000):
        rng = np.random.default_rng(seed)
        self.n_taxa, self.n_sites, self.n_states, self.n_cats = n_taxa, n_sites, n_states, n_cats
        self.seed, self.block_sites = seed, block_sites
        self.tree, self.root = random_tree(n_taxa, rng, bl_mean)
        if n_states == 2:
            self.pi = np.array([0.7, 0.3])
            self.er = np.array([1.0])
        else:
            self.pi = rng.dirichlet(np.full(n_states, 5.0))
            self.er = rng.dirichlet(np.ones(n_states * (n_states - 1) // 2))
        self.alpha = alpha
        from . import config
        saved = config.N_CATS
        config.N_CATS = n_cats
        try:
            self.rates = [float(r) for r in get_siterates(alpha)]
        finally:
            config.N_CATS = saved
        self._pm = None

    def edge_order(self):
        """Children-before-parents edge list in the reference's order (mcmc_gamma.pyx:200-218,587)."""
        kids = {}
        for (p, c) in self.tree:
            kids.setdefault(p, []).append(c)
        out, stack = [], [self.root]
        while stack:
            nd = stack.pop()
            x, y = kids[nd]
            out += [(nd, x), (nd, y)]
            if y > self.n_taxa:
                stack.append(y)
            if x > self.n_taxa:
                stack.append(x)
        return out[::-1]

    def codes(self, lo, hi):
        assert lo % self.block_sites == 0 and (hi % self.block_sites == 0 or hi == self.n_sites)
        if self._pm is None:
            self._pm = transition_matrices(fnGTR(self.er, self.pi), self.tree, self.rates)
        parts = []
        for b0 in range(lo, hi, self.block_sites):
            n = min(self.block_sites, hi - b0)
            parts.append(simulate_block(self.tree, self.root, self.n_taxa, self.pi, self._pm, self.n_cats, n,
                                        (self.seed, b0 // self.block_sites)))
        return np.concatenate(parts, axis=1)


def shard_bounds(n_sites, rank, world, granule):
    """Contiguous [lo, hi) column slice of `rank`, cut at multiples of `granule`; the last rank takes
    the remainder.  Site patterns are independent through the whole pruning pass
    (ML_gamma.pyx:24-38), so a shard needs nothing from its neighbours until the final sum."""
    n_gran = (n_sites + granule - 1) // granule
    per = n_gran // world
    extra = n_gran % world
    lo_g = rank * per + min(rank, extra)
    hi_g = lo_g + per + (1 if rank < extra else 0)
    return min(lo_g * granule, n_sites), min(hi_g * granule, n_sites)

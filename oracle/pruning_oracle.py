"""CPU oracle for the CyBayes tree-likelihood hot path -- TEST INFRASTRUCTURE ONLY.

A plain NumPy restatement of the reference algorithm, used as the checker by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg.  The
product path (``cybayes_b200/``) never imports this module and has no CPU fallback.

Parity pins: this restatement is checked (tests/test_oracle.py) against
 (a) the committed golden vectors in tests/golden/ that were produced by running the
     UNMODIFIED compiled reference (tests/golden/make_golden.py), and
 (b) the compiled reference itself (oracle/_ref/, built by oracle/build_ref.sh) when it
     is present.
The reference ships no tests or golden vectors of its own.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import linalg as _linalg
from scipy.special import gammainc as _gammainc
from scipy.stats import chi2 as _chi2

N_CATS = 4  # config.pyx:12

MISSING = ("?", "-")


# --------------------------------------------------------------------------- leaves
def encode_leaf(tokens, alphabet):
    """0/1 matrix (S, P) for one taxon -- utils.pyx:94-120 (sites2Mat).

    '?' and '-' are all-ones columns, 'a/b' is multi-hot, anything else one-hot at
    ``alphabet.index(token)``.
    """
    S = len(alphabet)
    pos = {a: i for i, a in enumerate(alphabet)}
    m = np.zeros((S, len(tokens)))
    for p, tok in enumerate(tokens):
        if tok in MISSING:
            m[:, p] = 1.0
        elif "/" in tok:
            for t in tok.split("/"):
                m[pos[t], p] = 1.0
        else:
            m[pos[tok], p] = 1.0
    return m


def read_phylip(path, reader):
    """Restates the three readers (utils.pyx:11-37, 39-65, 67-92).

    Returns (n_leaves, n_chars, alphabet, site_dict, ll_mats, taxa_list, n_sites) with
    ll_mats = {1-based taxon id: (S, P) float64}.
    """
    with open(path) as fh:
        n_leaves, n_sites = (int(x) for x in fh.readline().strip().split(" "))
        alphabet = ["0", "1"] if reader == "readBinaryPhy" else []
        rows, taxa = {}, []
        for line in fh:
            if len(line.strip()) < 1:
                continue
            if reader == "readPhy":
                name, vec = line.strip().split("\t")
                toks = vec.split(" ")
                seen = [t for tok in toks for t in tok.split("/")]
            else:
                name, vec = line.strip().split()
                toks = vec if reader == "readBinaryPhy" else list(vec)
                seen = list(vec)
            name = name.replace(" ", "")
            for t in seen:
                if t not in alphabet and t not in MISSING:
                    alphabet.append(t)
            rows[name] = toks
            taxa.append(name)
    ll = {taxa.index(k) + 1: encode_leaf(v, alphabet) for k, v in rows.items()}
    return n_leaves, len(alphabet), alphabet, rows, ll, taxa, n_sites


# ------------------------------------------------------------------------ traversal
def children_of(tree):
    """parent -> [children] in dict insertion order -- mcmc_gamma.pyx:220-232."""
    kids = {}
    for (p, c) in tree:
        kids.setdefault(p, []).append(c)
    return kids


def parent_of(tree):
    """child -> parent -- mcmc_gamma.pyx:234-242."""
    return {c: p for (p, c) in tree}


def preorder_edges(kids, node, n_taxa):
    """Recursive edge list, both edges of a node first, then the left and right
    subtrees -- mcmc_gamma.pyx:200-218.  Callers reverse it (``[::-1]``)."""
    x, y = kids[node]
    out = [(node, x), (node, y)]
    if x > n_taxa:
        out += preorder_edges(kids, x, n_taxa)
    if y > n_taxa:
        out += preorder_edges(kids, y, n_taxa)
    return out


def edge_order(tree, root, n_taxa):
    """The ``state['postorder']`` list: reversed pre-order -- mcmc_gamma.pyx:587."""
    return preorder_edges(children_of(tree), root, n_taxa)[::-1]


def path_to_root(parents, node, root):
    """Ancestors of ``node`` ending with the root -- mcmc_gamma.pyx:26-38."""
    out = []
    while True:
        node = parents[node]
        out.append(node)
        if node == root:
            return out


# ------------------------------------------------------------------- substitution
def site_rates(alpha, n_cats=N_CATS):
    """Mean-of-quantile discrete Gamma rates -- mcmc_gamma.pyx:596-602.

    ``alpha`` is a C float in the reference (fp32 truncation).
    """
    alpha = float(np.float32(alpha))
    cuts = [_chi2.isf(1 - p, 2 * alpha) for p in np.arange(1.0 / n_cats, 1, 1.0 / n_cats)]
    inc = [_gammainc(alpha + 1, c * alpha) for c in cuts]
    r = [inc[0] * n_cats]
    for i in range(1, n_cats - 1):
        r.append((inc[i] - inc[i - 1]) * n_cats)
    r.append((1.0 - inc[-1]) * n_cats)
    return r


def f81_beta(pi):
    """1 / (1 - pi.pi) -- mcmc_gamma.pyx:467,580."""
    pi = np.asarray(pi)
    return 1 / (1 - np.dot(pi, pi))


def gtr_q(er, pi):
    """Normalised GTR rate matrix -- mcmc_gamma.pyx:484-505."""
    pi = np.asarray(pi, dtype=float)
    S = pi.shape[0]
    R = np.zeros((S, S))
    R[np.triu_indices(S, 1)] = np.asarray(er, dtype=float)
    R = R + R.T
    Q = np.dot(R, np.diag(pi))
    Q += np.diag(-np.sum(Q, axis=-1))
    beta = -1.0 / np.dot(pi, np.diag(Q))
    return Q * beta


def p_matrix(model, binary, pi, er, beta, d, gtr_via="expm", Q=None):
    """One P(d) -- mcmc_gamma.pyx:372-401 / 450-482 with ptJC :507-514,
    binaryptF81 :516-525, ptF81 :527-547.  ``d`` is t * category rate."""
    pi = np.asarray(pi, dtype=float)
    S = pi.shape[0]
    if model == "JC":
        x = math.exp(-beta * d)
        y = (1.0 - x) / S
        P = np.full((S, S), y)
        np.fill_diagonal(P, x + y)
        return P
    if model == "F81":
        x = math.exp(-beta * d)
        y = 1.0 - x
        if binary:
            return np.array([[pi[0] + pi[1] * x, pi[1] * y], [pi[0] * y, pi[1] + pi[0] * x]])
        P = np.tile(pi * y, (S, 1))
        P[np.diag_indices(S)] = pi * y + x
        return P
    if model == "GTR":
        Q = gtr_q(er, pi) if Q is None else Q
        if gtr_via == "expm":
            return _linalg.expm(Q * d)
        lam, U, Uinv = gtr_eigensystem(Q, pi)
        return np.eye(S) + (U * np.expm1(lam * d)) @ Uinv
    raise ValueError(model)


def gtr_eigensystem(Q, pi):
    """Eigen-decomposition of a reversible Q through the symmetrised form
    diag(sqrt(pi)) Q diag(1/sqrt(pi)).  Not in the reference (it calls
    scipy.linalg.expm per edge, mcmc_gamma.pyx:481); used to validate the product's
    batched GTR builder."""
    s = np.sqrt(np.asarray(pi, dtype=float))
    Bm = (Q * s[:, None]) / s[None, :]
    Bm = 0.5 * (Bm + Bm.T)
    lam, V = np.linalg.eigh(Bm)
    return lam, V / s[:, None], V.T * s[None, :]


def prob_t(model, binary, pi, tree, er, mean_rate, beta=None, gtr_via="expm"):
    """dict edge -> P for one category rate -- mcmc_gamma.pyx:439-482.
    JC uses the caller's beta (config.NORM_BETA, set at :580); F81 recomputes it (:467)."""
    if model == "F81":
        beta = f81_beta(pi)
    if model == "GTR" and gtr_via == "expm":
        Q = gtr_q(er, pi)
        return {e: _linalg.expm(Q * t * mean_rate) for e, t in tree.items()}  # (Q*t)*r as at :481
    Q = gtr_q(er, pi) if model == "GTR" else None
    return {e: p_matrix(model, binary, pi, er, beta, t * mean_rate, gtr_via, Q) for e, t in tree.items()}


# ----------------------------------------------------------------------- likelihood
def mat_ml(pi, root, ll_mats, edges, tmats, n_sites, n_taxa, n_cats=N_CATS):
    """Full pruning pass, unscaled, exactly as ML_gamma.pyx:7-42."""
    ll = np.zeros(n_sites)
    caches = []
    for p_t in tmats:
        part = {}
        for parent, child in edges:
            src = ll_mats[child] if child <= n_taxa else part[child]
            v = p_t[parent, child].dot(src)
            if parent not in part:
                part[parent] = v
            else:
                part[parent] *= v
        ll += np.dot(pi, part[root]) / np.float32(n_cats)
        caches.append(part)
    with np.errstate(divide="ignore"):
        return np.sum(np.log(ll)), caches


def cache_mat_ml(pi, root, ll_mats, cache, nodes_recompute, edges, tmats, n_sites, n_taxa, n_cats=N_CATS):
    """Dirty-path pass as ML_gamma.pyx:83-118: parents in ``nodes_recompute`` are
    recomputed, every other parent is aliased from ``cache``."""
    ll = np.zeros(n_sites)
    caches = []
    dirty = set(nodes_recompute)
    for k, p_t in enumerate(tmats):
        part = {}
        for parent, child in edges:
            if parent in dirty:
                src = ll_mats[child] if child <= n_taxa else part[child]
                v = p_t[parent, child].dot(src)
                if parent not in part:
                    part[parent] = v
                else:
                    part[parent] *= v
            else:
                part[parent] = cache[k][parent]
        ll += np.dot(pi, part[root]) / (n_cats * 1.0)
        caches.append(part)
    with np.errstate(divide="ignore"):
        return np.sum(np.log(ll)), caches


def mat_ml_scaled(pi, root, ll_mats, edges, tmats, n_sites, n_taxa, n_cats=N_CATS, weights=None,
                  site_lnl=False, keep=None):
    """Same recursion with per-site power-of-two rescaling (exact in fp64), so that deep
    trees do not underflow (the reference has no rescaling).  After each
    completed node every site is divided by 2**e, e = exponent of the max over
    categories and states, and e is accumulated per site.  Where the unscaled
    recursion stays in range both give the same mantissas bit for bit.

    Returns (lnL, per-site lnL if asked).  `keep`: dict filled with node -> (mantissas (C, S, P),
    exponents (P,)) for the node ids it already holds as keys; every other partial is dropped as
    soon as its parent consumed it (each node has one parent), so memory stays near the tree width."""
    C = len(tmats)
    part, expo, nkids = {}, {}, {}
    for parent, child in edges:
        if child <= n_taxa:
            v = np.stack([tmats[k][parent, child].dot(ll_mats[child]) for k in range(C)])
            e = 0
        else:
            v = np.stack([tmats[k][parent, child].dot(part[child][k]) for k in range(C)])
            e = expo[child]
            if keep is not None and child in keep:
                keep[child] = (part[child], np.asarray(expo[child]))
            del part[child], expo[child]
        if parent not in part:
            part[parent], expo[parent], nkids[parent] = v, e, 1
        else:
            part[parent] = part[parent] * v
            expo[parent] = expo[parent] + e
            nkids[parent] += 1
        if nkids[parent] == 2 and parent != root:
            m = part[parent].max(axis=(0, 1))
            _, ex = np.frexp(m)
            ex = np.where(m > 0, ex - 1, 0)
            part[parent] = np.ldexp(part[parent], -ex[None, None, :])
            expo[parent] = expo[parent] + ex
    ll = np.zeros(n_sites)
    for k in range(C):
        ll += np.dot(pi, part[root][k]) / (n_cats * 1.0)
    with np.errstate(divide="ignore"):
        per_site = np.log(ll) + np.asarray(expo[root]) * math.log(2.0)
    if weights is not None:
        per_site = per_site * weights
    total = float(np.sum(per_site))
    return (total, per_site) if site_lnl else total


def leaves_from_codes(codes, n_states, amb_sets=None):
    """{taxon id: (S, P) 0/1 matrix} from a state-code matrix (code < S: one-hot; S + k: row k of
    amb_sets, row 0 = all ones) -- the columns utils.pyx:94-120 would produce for those cells."""
    amb = np.ones((1, n_states)) if amb_sets is None else np.asarray(amb_sets, dtype=float)
    table = np.vstack([np.eye(n_states), amb])
    return {t + 1: np.ascontiguousarray(table[codes[t].astype(np.int64)].T) for t in range(codes.shape[0])}


def mat_ml_scaled_codes(pi, root, codes, n_states, amb_sets, edges, tmats, n_taxa, n_cats=N_CATS, chunk=16384,
                        keep_nodes=()):
    """mat_ml_scaled over a big state-code matrix, one chunk of columns at a time (sites are independent
    through the whole pass, ML_gamma.pyx:24-38, so the per-site values are those of a single pass).
    Returns (lnL, per-site lnL, {node: (mantissas (C, S, P), exponents (P,))} for keep_nodes)."""
    n_sites = codes.shape[1]
    per_site = np.empty(n_sites)
    kept = {n: ([], []) for n in keep_nodes}
    for lo in range(0, n_sites, chunk):
        hi = min(n_sites, lo + chunk)
        keep = {n: None for n in keep_nodes}
        _, ps = mat_ml_scaled(pi, root, leaves_from_codes(codes[:, lo:hi], n_states, amb_sets), edges, tmats,
                              hi - lo, n_taxa, n_cats, site_lnl=True, keep=keep)
        per_site[lo:hi] = ps
        for n in keep_nodes:
            kept[n][0].append(keep[n][0])
            kept[n][1].append(np.broadcast_to(keep[n][1], (hi - lo,)))
    out = {n: (np.concatenate(m, axis=2), np.concatenate(e)) for n, (m, e) in kept.items()}
    return float(np.sum(per_site)), per_site, out

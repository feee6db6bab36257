"""Substitution models: transition-probability matrices P(t r) for every branch x discrete-Gamma
category, built on the GPU in one batched launch, and the Gamma category rates.

Restates the live substitution code of the reference (mcmc_gamma.pyx:372-401, 439-547,
596-602; mcmc.pyx:354-434).  ``get_prob_t`` / ``get_edge_transition_mat`` keep their
signatures but return *device handles*: a ``PMatTable`` behaves like the reference's
``{(parent, child): ndarray}`` dict (get / set / del by edge, values with ``.copy()``), the
matrices themselves live in the engine's slot pool and are produced by ``cb_pmat_build``
when the next likelihood evaluation flushes the queue.
"""
from __future__ import annotations

import math
import os

import numpy as np
from scipy.special import chdtri, gammainc
from scipy.stats import chi2

from . import config
from ._lib import CB_MODEL_F81, CB_MODEL_F81_BINARY, CB_MODEL_GTR_EIG, CB_MODEL_JC

# exp(-beta d) for JC/F81 is taken from the host libm (the reference's c_exp) for up to this many
# matrices per call, which makes those matrices bit-identical to the reference's; larger batches
# (the synthetic 1024-taxon configs) use the device exp (<= 1 ulp).  0 = always device.
HOST_EXP_MAX = int(os.environ.get("CYBAYES_HOST_EXP_MAX", "4096"))

_engine_hook = None  # set by likelihood.py: (n_cats) -> Engine holding config.LEAF_LLMAT


def _engine(n_cats=None):
    if _engine_hook is None:
        from . import likelihood  # noqa: F401  (installs the hook)
    return _engine_hook(n_cats)


class PMatrix:
    """One S x S transition matrix resident on the device.  Immutable, so ``copy`` is the
    identity (the driver copies to restore after a rejected move, mat_mcmc_gamma.py:146,206)."""
    __slots__ = ("engine", "slot", "owner")

    def __init__(self, engine, slot, owner):
        self.engine, self.slot, self.owner = engine, slot, owner

    def copy(self):
        return self

    __copy__ = copy

    def __deepcopy__(self, memo):
        return self

    @property
    def shape(self):
        return (self.engine.n_states, self.engine.n_states)

    def __array__(self, dtype=None, copy=None):
        a = self.engine.download_pmats(np.array([self.slot], dtype=np.int32))[0]
        return a if dtype is None else a.astype(dtype)

    def dot(self, other):
        return np.asarray(self).dot(other)


class PMatTable:
    """Mapping edge -> PMatrix over device slots (one rate category).

    A table as get_prob_t / get_prob_t_all return it is a run of consecutive slots in the order of the edge list it
    was built from; the edge -> slot dict is only materialised when somebody indexes or edits the table (`_slots`).
    Until then `pristine()` hands matML the (edge list, first slot) pair, and a full evaluation maps the run onto
    its op order with one cached permutation instead of one dict lookup per edge."""
    __slots__ = ("engine", "_map", "_owners", "_block", "_edges", "_base")

    def __init__(self, engine, edges, block, base=None, n=None):
        # (base, n): this table's slice of a block shared by the tables of all rate categories
        self.engine = engine
        self._block = block
        if base is None:
            base, n = block.base, block.n
        if len(edges) != n:
            raise ValueError("one slot per edge")
        self._edges, self._base = edges, base
        self._map = None
        self._owners = {}

    @property
    def _slots(self):
        m = self._map
        if m is None:
            m = self._map = dict(zip(self._edges, range(self._base, self._base + len(self._edges))))
            self._edges = None
        return m

    def pristine(self):
        """(edge list, first slot) while nobody has looked inside the table, else None."""
        return (self._edges, self._base) if self._map is None else None

    def __getitem__(self, edge):
        return PMatrix(self.engine, self._slots[edge], self._owners.get(edge, self._block))

    def __setitem__(self, edge, value):
        if not isinstance(value, PMatrix):
            value = host_matrix(self.engine, value)
        if value.engine is not self.engine:
            raise ValueError("transition matrix belongs to a different alignment")
        self._slots[edge] = value.slot
        if value.owner is self._block:
            self._owners.pop(edge, None)
        else:
            self._owners[edge] = value.owner  # keeps the foreign slot block alive

    def __delitem__(self, edge):
        del self._slots[edge]
        self._owners.pop(edge, None)

    def __contains__(self, edge):
        return edge in self._slots

    def __len__(self):
        return len(self._slots)

    def __iter__(self):
        return iter(self._slots)

    def keys(self):
        return self._slots.keys()

    def items(self):
        return ((e, self[e]) for e in self._slots)

    def values(self):
        return (self[e] for e in self._slots)

    def get(self, edge, default=None):
        return self[edge] if edge in self._slots else default

    def copy(self):
        t = PMatTable.__new__(PMatTable)
        t.engine, t._block = self.engine, self._block
        t._edges, t._base = self._edges, self._base
        t._map = None if self._map is None else dict(self._map)
        t._owners = dict(self._owners)
        return t


def host_matrix(engine, mat):
    """Wrap a host ndarray (e.g. a reference-style P matrix) as a device PMatrix."""
    block = engine.alloc_slots(1)
    engine.upload_pmats(np.array([block.base], dtype=np.int32), np.asarray(mat, dtype=np.float64))
    return PMatrix(engine, block.base, block)


def table_from_host(engine, mapping):
    """Upload a reference-style ``{edge: ndarray}`` dict (host buffers) into a PMatTable."""
    edges = list(mapping.keys())
    block = engine.alloc_slots(len(edges))
    vals = list(mapping.values())
    try:   # fastest way to gather thousands of small (S, S) arrays into one buffer
        mats = np.concatenate(vals)
        if mats.dtype != np.float64 or mats.shape != (len(vals) * engine.n_states, engine.n_states):
            raise ValueError
    except (ValueError, TypeError):
        mats = np.array([np.asarray(v, dtype=np.float64) for v in vals])
    engine.upload_pmats(np.arange(block.base, block.base + block.n, dtype=np.int32), mats)
    return PMatTable(engine, edges, block)


# ----------------------------------------------------------------------------- models
def f81_beta(pi):
    return 1 / (1 - np.dot(pi, pi))


def fnGTR(er, pi):
    """Normalised reversible rate matrix (mcmc_gamma.pyx:484-505)."""
    pi = np.asarray(pi, dtype=np.float64)
    er = np.asarray(er, dtype=np.float64)
    n_states = pi.shape[0]
    R = np.zeros((n_states, n_states))
    R[np.triu_indices(n_states, 1)] = er
    R = R + R.T
    Q = np.dot(R, np.diag(pi))
    Q += np.diag(-np.sum(Q, axis=-1))
    beta = -1.0 / np.dot(pi, np.diag(Q))
    return Q * beta


_gtr_cache = {"key": None, "eig": None}


def gtr_eigensystem(pi, er):
    """[lambda | U | U^-1] of Q through the symmetric form diag(sqrt pi) Q diag(1/sqrt pi); one
    decomposition per (pi, rates), shared by all branches and categories (the reference calls
    scipy.linalg.expm per branch, mcmc_gamma.pyx:481)."""
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    er = np.ascontiguousarray(er, dtype=np.float64)
    key = (pi.tobytes(), er.tobytes())
    if _gtr_cache["key"] == key:
        return _gtr_cache["eig"]
    Q = fnGTR(er, pi)
    s = np.sqrt(pi)
    B = (Q * s[:, None]) / s[None, :]
    B = 0.5 * (B + B.T)
    lam, V = np.linalg.eigh(B)
    U = V / s[:, None]
    Uinv = V.T * s[None, :]
    eig = np.ascontiguousarray(np.concatenate([lam, U.ravel(), Uinv.ravel()]))
    _gtr_cache["key"], _gtr_cache["eig"] = key, eig
    return eig


def _queue(engine, model_name, binary, pi, rates, slots, d, norm_beta):
    """Queue the build of P(d[i]) into slots[i] for the named model."""
    n = len(d)
    pi_a = np.array(pi, dtype=np.float64)
    if model_name == "GTR":
        engine.queue_build(CB_MODEL_GTR_EIG, pi_a, 0.0, gtr_eigensystem(pi_a, rates), slots, d, None)
        return
    beta = float(norm_beta)
    x = None
    if 0 < n <= HOST_EXP_MAX:
        exp = math.exp
        nb = -beta
        x = [exp(nb * v) for v in (d.tolist() if isinstance(d, np.ndarray) else d)]
    if model_name == "JC":
        engine.queue_build(CB_MODEL_JC, pi_a, beta, None, slots, d, x)
    elif model_name == "F81":
        engine.queue_build(CB_MODEL_F81_BINARY if binary else CB_MODEL_F81, pi_a, beta, None, slots, d, x)
    else:
        raise ValueError(f"unknown model {model_name!r}")


def get_prob_t(pi, edges_dict, rates, mean_rate, n_cats=None):
    """P(t * mean_rate) for every edge of the tree, one rate category (mcmc_gamma.pyx:439-482).
    F81 refreshes config.NORM_BETA = 1/(1 - pi.pi) (:467); JC uses the stored value (:457).
    `n_cats` selects the device context (default config.N_CATS; 1 for the mcmc.pyx surface)."""
    engine = _engine(n_cats)
    model = config.MODEL
    if model == "F81":
        config.NORM_BETA = f81_beta(np.asarray(pi))
    edges = list(edges_dict)
    block = engine.alloc_slots(len(edges))
    d = np.array([v * mean_rate for v in edges_dict.values()], dtype=np.float64)
    slots = np.arange(block.base, block.base + block.n, dtype=np.int32)
    _queue(engine, model, config.IN_DTYPE == "bin", pi, rates, slots, d, config.NORM_BETA)
    return PMatTable(engine, edges, block)


def get_prob_t_all(pi, edges_dict, rates, site_rates, n_cats=None):
    """[get_prob_t(pi, edges_dict, rates, r) for r in site_rates] as one queue entry over one slot block: the same
    matrices (same host exp, same device arithmetic), a quarter of the host bookkeeping.  Used by the restated
    driver for full-pass proposals; the unchanged reference driver calls get_prob_t per category."""
    engine = _engine(n_cats)
    model = config.MODEL
    if model == "F81":
        config.NORM_BETA = f81_beta(np.asarray(pi))
    edges = list(edges_dict)
    n_e, n_c = len(edges), len(site_rates)
    block = engine.alloc_slots(n_e * n_c)
    # d[k, e] = t_e * r_k, the product the reference forms per edge (mcmc_gamma.pyx:455,469,481): one rounding each
    d = np.multiply.outer(np.array(site_rates, dtype=np.float64), np.fromiter(edges_dict.values(), dtype=np.float64,
                                                                           count=n_e)).ravel()
    slots = np.arange(block.base, block.base + block.n, dtype=np.int32)
    _queue(engine, model, config.IN_DTYPE == "bin", pi, rates, slots, d, config.NORM_BETA)
    return [PMatTable(engine, edges, block, block.base + k * n_e, n_e) for k in range(n_c)]


def get_edge_transition_mat(pi, rates, d, n_cats=None):
    """One P(d) for a single branch move (mcmc_gamma.pyx:372-401); d = t * category rate."""
    engine = _engine(n_cats)
    model = config.MODEL
    if model == "F81":
        config.NORM_BETA = f81_beta(np.asarray(pi))
    block = engine.alloc_slots(1)
    _queue(engine, model, config.IN_DTYPE == "bin", pi, rates, [block.base], [float(d)], config.NORM_BETA)
    return PMatrix(engine, block.base, block)


def get_edge_transition_mats(pi, rates, ds, n_cats=None):
    """[P(d) for d in ds] in one queue entry (one slot block): what a branch move needs for all rate
    categories (and both branches of a node slide).  Same matrices as get_edge_transition_mat."""
    engine = _engine(n_cats)
    model = config.MODEL
    if model == "F81":
        config.NORM_BETA = f81_beta(np.asarray(pi))
    n = len(ds)
    block = engine.alloc_slots(n)
    base = block.base
    _queue(engine, model, config.IN_DTYPE == "bin", pi, rates, list(range(base, base + n)), [float(d) for d in ds],
           config.NORM_BETA)
    return [PMatrix(engine, base + i, block) for i in range(n)]


_points = {}


def _quantile_points(n):
    pts = _points.get(n)
    if pts is None:
        pts = _points[n] = list(np.arange(1.0 / n, 1, 1.0 / n))
    return pts


def get_siterates(alpha):
    """Mean rates of the N_CATS equiprobable discrete-Gamma categories (mcmc_gamma.pyx:596-602).
    `alpha` is a C float in the reference: it is rounded to fp32 first."""
    alpha = float(np.float32(alpha))
    n = config.N_CATS
    # chi2.isf(q, df) is scipy.special.chdtri(df, q) behind ~100 us of argument checking per call: the direct
    # call returns the same bits (checked over 3 500 alphas in tests/test_host_logic.py)
    cutoffs = [chdtri(2 * alpha, 1 - p) for p in _quantile_points(n)]
    cum = [gammainc(alpha + 1, c * alpha) for c in cutoffs]
    site_rates = [cum[0] * n]
    for i in range(1, n - 1):
        site_rates.append((cum[i] - cum[i - 1]) * n)
    site_rates.append((1.0 - cum[-1]) * n)
    return site_rates

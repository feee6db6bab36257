"""TEST INFRASTRUCTURE: a stand-in for cybayes_b200.engine.Engine that evaluates op lists with
the NumPy oracle.  It lets the CPU-only suite exercise the host logic (readers, pattern
compression, op-list planning, P-slot bookkeeping, snapshot lifetimes, the MCMC driver's random
number consumption) without a GPU.  It is never importable from the product package."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pruning_oracle as oracle  # noqa: E402

from cybayes_b200._lib import CB_MODEL_F81, CB_MODEL_F81_BINARY, CB_MODEL_GTR_EIG, CB_MODEL_JC  # noqa: E402
from cybayes_b200.engine import SlotPool  # noqa: E402


class FakeEngine(SlotPool):
    instances = []

    def __init__(self, codes, n_states, n_cats, amb_sets=None, weights=None, device=None):
        self._init_slots()
        self.n_taxa, self.n_patterns = codes.shape
        self.n_states, self.n_cats = int(n_states), int(n_cats)
        amb = np.ones((1, n_states)) if amb_sets is None else np.asarray(amb_sets, dtype=float)
        table = np.vstack([np.eye(n_states), amb])
        self.leaves = {t + 1: np.ascontiguousarray(table[codes[t].astype(np.int64)].T) for t in range(self.n_taxa)}
        self.weights = np.ones(self.n_patterns) if weights is None else np.asarray(weights, dtype=float)
        self.pm = {}
        self.snaps = {}
        self._next_snap = 0
        self.n_evals = 0
        self.n_ops = 0
        self.gtr_via = "expm"
        FakeEngine.instances.append(self)

    def _reserve(self, n):
        pass

    def close(self):
        pass

    # P matrices -----------------------------------------------------------------
    def upload_pmats(self, slots, mats):
        mats = np.asarray(mats, dtype=float).reshape(len(slots), self.n_states, self.n_states)
        for s, m in zip(np.asarray(slots).tolist(), mats):
            self.pm[s] = m.copy()

    def download_pmats(self, slots):
        return np.stack([self.pm[s] for s in np.asarray(slots).tolist()])

    def queue_build(self, model, pi, beta, gtr, slots, d, x=None):
        name = {CB_MODEL_JC: "JC", CB_MODEL_F81: "F81", CB_MODEL_F81_BINARY: "F81", CB_MODEL_GTR_EIG: "GTR"}[model]
        if name == "GTR":
            S = self.n_states
            lam, U, Uinv = gtr[:S], gtr[S:S + S * S].reshape(S, S), gtr[S + S * S:].reshape(S, S)
            Q = (U * lam) @ Uinv  # the rate matrix back from its eigensystem
        for s, dd in zip(np.asarray(slots).tolist(), np.asarray(d, dtype=float).tolist()):
            if name == "GTR":
                self.pm[s] = oracle._linalg.expm(Q * dd) if self.gtr_via == "expm" else (U * np.exp(lam * dd)) @ Uinv
            else:
                self.pm[s] = oracle.p_matrix(name, model == CB_MODEL_F81_BINARY, pi, None, beta, dd)

    def flush_builds(self):
        pass

    # evaluation -------------------------------------------------------------------
    def _run(self, snapshot, nodes, children, pslots, pi):
        base = self.snaps[snapshot] if snapshot is not None else {}
        new = {}
        C = self.n_cats
        for i, node in enumerate(np.asarray(nodes).tolist()):
            acc = None
            for kx in range(2):
                ch = int(children[2 * i + kx])
                if ch <= self.n_taxa:
                    src = [self.leaves[ch]] * C
                else:
                    src = new[ch] if ch in new else base[ch]
                v = np.stack([self.pm[int(pslots[2 * i + kx, c])].dot(src[c]) for c in range(C)])
                acc = v if acc is None else acc * v
            new[node] = acc
            self.n_ops += 1
        root = int(nodes[-1])
        ll = np.zeros(self.n_patterns)
        for c in range(C):
            ll += np.dot(pi, new[root][c]) / (C * 1.0)
        with np.errstate(divide="ignore"):
            lnl = float(np.sum(self.weights * np.log(ll)))
        return lnl, base, new

    def eval(self, snapshot, nodes, children, pslots, pi, want_snapshot=True, store_root=False,
             force_levels=False, sync=True, force_walk=False, no_fold=False):
        self.n_evals += 1
        lnl, base, new = self._run(snapshot, nodes, children, pslots, np.asarray(pi, dtype=float))
        sid = -1
        if want_snapshot:
            merged = dict(base)
            merged.update(new)
            sid = self._next_snap
            self._next_snap += 1
            self.snaps[sid] = merged
        return lnl, sid

    def eval_batch(self, snapshot, offsets, nodes, children, pslots, pi):
        out = []
        for b in range(len(offsets) - 1):
            lo, hi = int(offsets[b]), int(offsets[b + 1])
            out.append(self._run(snapshot, nodes[lo:hi], children[2 * lo:2 * hi], pslots[2 * lo:2 * hi],
                                 np.asarray(pi, dtype=float))[0])
        return np.array(out)

    def release_snapshot(self, snap):
        self.snaps.pop(snap, None)

    def read_partial(self, snap, node, with_scale=False):
        a = self.snaps[snap][node]
        return (a, np.zeros(self.n_patterns, dtype=np.int32)) if with_scale else a

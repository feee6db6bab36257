"""Measurement aid: time the multistate kernels (DMMA vs the plain FP64 kernel) on synthetic data.
usage: python profiles/microbench/multistate_probe.py S N_TAXA N_SITES"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from cybayes_b200 import _lib
from cybayes_b200.engine import Engine
from cybayes_b200.likelihood import _Plan
from cybayes_b200.subst import gtr_eigensystem
from cybayes_b200.synthetic import SyntheticAlignment

S, N, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
aln = SyntheticAlignment(N, P, S, 20260102, block_sites=P)
t0 = time.time(); codes = aln.codes(0, P); print("generated", codes.shape, f"{time.time()-t0:.1f}s", flush=True)
C = 4
eng = Engine(codes, S, C)
plan = _Plan(aln.edge_order())
ekeys = list(aln.tree.keys()); n_e = len(ekeys)
block = eng.alloc_slots(n_e * C)
slots = np.arange(block.base, block.base + n_e * C, dtype=np.int32)
d = np.array([aln.tree[e] * r for r in aln.rates for e in ekeys])
eng.queue_build(_lib.CB_MODEL_GTR_EIG, aln.pi, 0.0, gtr_eigensystem(aln.pi, aln.er), slots, d)
slot_of = {(k, e): block.base + k * n_e + i for k in range(C) for i, e in enumerate(ekeys)}
pslots = np.array([[slot_of[k, e] for k in range(C)] for e in plan.edge_keys], dtype=np.int32)
n_int_edges = sum(1 for (p, c) in ekeys if c > N)
flops = C * P * (2.0 * S * S * n_int_edges + S * (N - 1))
for snap in (False, True):
    for force_levels in (False, True):
        ms = []
        for it in range(4):
            lnl, sn = eng.eval(None, plan.nodes, plan.children, pslots, aln.pi, want_snapshot=snap, force_levels=force_levels)
            ms.append(eng.last_eval_ms())
            if sn >= 0: eng.release_snapshot(sn)
        m = min(ms[1:])
        print(f"S={S} N={N} P={P} snapshot={snap} levels={force_levels}: {m:.3f} ms  {flops/m/1e9:.2f} TFLOP/s  lnL={lnl!r}", flush=True)

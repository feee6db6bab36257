"""GPU parity at the benchmark shapes (-m gpu): the kernel instantiations that produce the headline numbers,
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
import pytest

import large_cases
from cybayes_b200 import likelihood

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_compression(monkeypatch):
    # the kernels must see every simulated column (the bench does the same): constant columns repeat
    monkeypatch.setattr(likelihood, "COMPRESS_MAX_SITES", 0)


def test_c4_shape_kernel_instantiations_vs_oracle(gpu_backend):
    """prune_s2_kernel<4,1,256,3>: single-launch walk, split walk, likelihood-only, dirty path, cached partials."""
    large_cases.run_c4_shape(1024, 131072, 16384)


def test_c5_shape_register_carried_dmma_vs_oracle(gpu_backend):
    """prune_dmma_rc_kernel<64, true> on a 256-taxon x 64-state x 8192-site simulated alignment."""
    large_cases.run_c5_shape(256, 8192, 64, 1024)

"""cybayes_b200 -- B200-native Felsenstein-pruning likelihood behind the CyBayes call surface.

Layout (only what the hot path needs):
  csrc/        hand-written sm_100a CUDA kernels + the C-ABI library (include/cybayes_b200.h)
  _lib.py      ctypes binding of that C ABI (no torch, no CPU fallback)
  engine.py    device context: leaf codes, P-matrix slots, partial-cache snapshots, evaluation
  alignment.py Phylip readers, leaf encoding, site-pattern compression   (utils.pyx)
  tree.py      traversal indexing                                         (mcmc_gamma.pyx:26-38,200-242)
  subst.py     P(t) builders and discrete-Gamma rates                     (mcmc_gamma.pyx:372-547,596-602)
  moves.py     proposals and start state                                  (mcmc_gamma.pyx:40-198,244-332,573-593)
  likelihood.py  matML / cache_matML surfaces                             (ML_gamma.pyx, ML.pyx)
  driver.py    Metropolis-Hastings loop                                   (mat_mcmc_gamma.py, mat_mcmc.py)
  compat/      top-level modules `config utils mcmc_gamma ML_gamma mcmc ML` so the reference's
               unchanged driver scripts run on top of this package
"""
__version__ = "0.1.0"

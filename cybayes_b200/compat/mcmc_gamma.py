"""Top-level alias so the reference's unchanged driver (`from mcmc_gamma import *`) resolves to
cybayes_b200.mcmc_gamma when this directory is on PYTHONPATH.  Like the reference module
(mcmc_gamma.pyx:6-7) importing it seeds both random generators with 1234."""
import importlib
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
np.random.seed(1234)
random.seed(1234)
sys.modules[__name__] = importlib.import_module("cybayes_b200.mcmc_gamma")

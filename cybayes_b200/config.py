"""Module-level blackboard shared by the driver and the library, as in the reference
(config.pyx:1-22): plain assignable attributes, N_CATS = 4."""
N_CHARS = 0
N_TAXA = 0
N_SITES = 0
N_GEN = 0
THIN = 0
N_CATS = 4
N_NODES = 0

ALPHABET = []
TAXA = []

NORM_BETA = 0.0

LEAF_LLMAT = {}

MODEL = ""
IN_DTYPE = ""

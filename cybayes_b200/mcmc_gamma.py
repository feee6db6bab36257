"""Mirror of the reference module ``mcmc_gamma`` (mcmc_gamma.pyx): same public names, so
``from mcmc_gamma import *`` gives the driver everything it uses, including ``np`` and
``random``."""
import copy  # noqa: F401  (exported by the reference module)
import random  # noqa: F401

import numpy as np  # noqa: F401
from scipy import linalg  # noqa: F401
from scipy.special import gammainc  # noqa: F401
from scipy.stats import chi2, dirichlet  # noqa: F401

from . import config
from .moves import (bl_exp_scale, epsilon, externalSPR, init_alpha_rate, init_pi_er, init_tree,  # noqa: F401
                    mvDualSlider, newick2bl, node_slider, rooted_NNI, rtree, scale_alpha, scale_edge,
                    scaler_alpha)
from .subst import fnGTR, get_edge_transition_mat, get_prob_t, get_siterates  # noqa: F401
from .tree import (adjlist2newickBL, adjlist2nodes_dict, adjlist2reverse_nodes_dict, get_path2root,  # noqa: F401
                   postorder)


def state_init():
    """Start state of a chain (mcmc_gamma.pyx:573-593): pi/rates, random tree, Gamma shape,
    edge order, and the P matrices of all branches for the N_CATS category rates."""
    state = {}
    pi, er = init_pi_er()
    config.NORM_BETA = 1 / (1 - np.dot(pi, pi))
    state["pi"] = pi
    state["rates"] = er
    state["tree"], state["root"] = init_tree()
    state["srates"] = init_alpha_rate()
    state["postorder"] = postorder(adjlist2nodes_dict(state["tree"]), state["root"])[::-1]
    state["transitionMat"] = [get_prob_t(state["pi"], state["tree"], state["rates"], r)
                              for r in get_siterates(state["srates"])]
    return state

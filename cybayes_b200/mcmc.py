"""Mirror of the reference module ``mcmc`` (mcmc.pyx), the single-rate twin of ``mcmc_gamma``.

Differences from mcmc_gamma that the reference has and that are kept:
no scale_alpha / GTR (node_slider exists, mcmc.pyx:59-89, same as mcmc_gamma's; the non-Gamma driver just never
proposes it); rooted_NNI returns four values (mcmc.pyx:124);
mvDualSlider draws ``random.uniform(epsilon, sum)`` over ``range(config.N_CHARS)`` (:179-181);
get_prob_t(pi, edges_dict, rates=None) has no rate argument (:427); get_edge_transition_mat
mutates and returns the dict it is given (:354-378); state_init builds a single P dict.
"""
import random  # noqa: F401

import numpy as np  # noqa: F401
from scipy.stats import dirichlet  # noqa: F401

from . import config, moves, subst
from .ML import cache_matML, matML  # noqa: F401
from .moves import (bl_exp_scale, epsilon, externalSPR, init_pi_er, init_tree, newick2bl, node_slider,  # noqa: F401
                    rtree, scale_edge, scaler_alpha)
from .tree import (adjlist2newickBL, adjlist2nodes_dict, adjlist2reverse_nodes_dict, get_path2root,  # noqa: F401
                   postorder)


def rooted_NNI(temp_edges_list, root_node):
    tree, new_postorder, hr, nodes_recompute, _ = moves.rooted_NNI(temp_edges_list, root_node)
    return tree, new_postorder, hr, nodes_recompute


def mvDualSlider(pi):
    i, j = random.sample(range(config.N_CHARS), 2)
    total = pi[i] + pi[j]
    x = random.uniform(epsilon, total)
    pi[i], pi[j] = x, total - x
    return pi, 0.0


def get_prob_t(pi, edges_dict, rates=None):
    if config.MODEL not in ("F81", "JC"):
        return None  # the reference falls off the end of the if-chain (mcmc.pyx:427-434)
    return subst.get_prob_t(pi, edges_dict, rates, 1.0, n_cats=1)


def get_edge_transition_mat(pi, rates, d, transition_mat, change_edge):
    if config.MODEL in ("F81", "JC"):
        transition_mat[change_edge] = subst.get_edge_transition_mat(pi, rates, d, n_cats=1)
    return transition_mat


def state_init():
    state = {}
    pi, er = init_pi_er()
    config.NORM_BETA = 1 / (1 - np.dot(pi, pi))
    state["pi"] = pi
    state["rates"] = er
    state["tree"], state["root"] = init_tree()
    state["postorder"] = postorder(adjlist2nodes_dict(state["tree"]), state["root"])[::-1]
    state["transitionMat"] = get_prob_t(state["pi"], state["tree"], state["rates"])
    return state

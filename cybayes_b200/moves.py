"""MCMC proposals and the start state (host side; the GPU never sees random numbers).

Restates mcmc_gamma.pyx:40-198 (moves), :244-332 (start tree, priors' draws) and :573-593
(state_init), and the differences of the non-Gamma twin mcmc.pyx.  What is pinned is the
*trace*: for a fixed seed the sequence of draws from Python's ``random`` and NumPy's legacy
global generator, and the insertion order of the tree dict (it feeds
``random.choice(list(tree))``), are the reference's, so a chain driven by these
functions accepts and rejects exactly like the reference chain.
"""
from __future__ import annotations

import math
import random

import numpy as np

from . import config
from .tree import adjlist2nodes_dict, adjlist2reverse_nodes_dict, get_path2root, postorder

bl_exp_scale = 0.1   # mean of the exponential branch-length prior (mcmc_gamma.pyx:21)
scaler_alpha = 1.0   # window of the multiplier proposals (mcmc_gamma.pyx:22)
epsilon = 1e-10      # mcmc_gamma.pyx:23


def _multiplier():
    """log c uniform on (-1/2, 1/2) * scaler_alpha, and c."""
    log_c = scaler_alpha * (random.random() - 0.5)
    return log_c, math.exp(log_c)


def scale_edge(temp_edges_dict):
    """Multiply one random branch by c (mcmc_gamma.pyx:40-60).
    Returns (tree, log c [Hastings], log prior ratio, edge)."""
    edge = random.choice(list(temp_edges_dict))
    old = temp_edges_dict[edge]
    log_c, c = _multiplier()
    new = old * c
    temp_edges_dict[edge] = new
    prior_ratio = -(new - old) / bl_exp_scale
    return temp_edges_dict, log_c, prior_ratio, edge


def node_slider(temp_edges_dict, root_node):
    """Rescale the two branches around an internal node and slide the node along their sum
    (mcmc_gamma.pyx:62-92).  Returns (tree, log c, log prior ratio, lower edge, upper edge)."""
    parent_of = adjlist2reverse_nodes_dict(temp_edges_dict)
    while True:
        edge = random.choice(list(temp_edges_dict))
        if edge[0] != root_node:
            break
    upper = (parent_of[edge[0]], edge[0])
    total = temp_edges_dict[upper] + temp_edges_dict[edge]
    log_c, c = _multiplier()
    new_total = total * c
    temp_edges_dict[upper] = new_total * random.random()
    temp_edges_dict[edge] = new_total - temp_edges_dict[upper]
    prior_ratio = -(new_total - total) / bl_exp_scale
    return temp_edges_dict, log_c, prior_ratio, edge, upper


def scale_alpha(alpha):
    """Multiplier move on the Gamma shape; `alpha` is a C float in the reference, i.e. rounded
    to fp32 on entry (mcmc_gamma.pyx:94-99)."""
    alpha = float(np.float32(alpha))
    log_c, c = _multiplier()
    new_alpha = alpha * c
    return new_alpha, log_c, -(new_alpha - alpha)


def _shuffle(x, getrandbits=random.getrandbits):
    """random.shuffle(x), draw for draw (Fisher-Yates from the top; randbelow by rejection on bit_length bits), without
    a Python-level call per element -- the shuffle of the edge list is a third of the cost of an NNI proposal."""
    for i in range(len(x) - 1, 0, -1):
        n = i + 1
        k = n.bit_length()
        r = getrandbits(k)
        while r >= n:
            r = getrandbits(k)
        x[i], x[r] = x[r], x[i]


def _nni(tree, root_node):
    kids = adjlist2nodes_dict(tree)
    order = list(tree.keys())
    _shuffle(order)
    for a, b in order:
        if b > config.N_TAXA:
            break
    sib_a, kids_b = kids[a], kids[b]
    src = sib_a[1] if sib_a[0] == b else sib_a[0]
    tgt = random.choice(kids_b)
    src_bl, tgt_bl = tree[a, src], tree[b, tgt]
    del tree[a, src], tree[b, tgt]
    tree[a, tgt] = tgt_bl
    tree[b, src] = src_bl
    # children lists of the new tree (= adjlist2nodes_dict(tree)): the two re-inserted edges are last in the dict
    kids[a] = [c for c in sib_a if c != src] + [tgt]
    kids[b] = [c for c in kids_b if c != tgt] + [src]
    new_postorder = postorder(kids, root_node)[::-1]
    nodes_recompute = [b] + get_path2root(adjlist2reverse_nodes_dict(tree), b, root_node)
    return tree, new_postorder, 0.0, nodes_recompute, [a, b, src, tgt]


def rooted_NNI(temp_edges_list, root_node):
    """Nearest-neighbour interchange across the first internal edge (a, b) of a shuffled edge
    list: b's child `tgt` swaps with b's sibling `src` (mcmc_gamma.pyx:101-134).
    Returns (tree, postorder, 0.0, dirty nodes [b .. root], [a, b, src, tgt])."""
    return _nni(temp_edges_list, root_node)


def externalSPR(edges_list, root_node):
    """Prune a random leaf (with its parent node) and regraft it on a random edge
    (mcmc_gamma.pyx:136-185).  The 'Hastings ratio' returned is r/(x+y) itself, not its log,
    and the driver adds it to the log ratio -- reproduced as is."""
    parent_of = adjlist2reverse_nodes_dict(edges_list)
    kids = adjlist2nodes_dict(edges_list)
    leaf = random.randint(1, config.N_TAXA)
    hub = parent_of[leaf]
    tgt = random.choice(list(edges_list))
    hastings_ratio = 0.0
    if not (hub == root_node or hub in tgt or parent_of[hub] in tgt):
        up = parent_of[hub]
        pair = kids[hub]
        other = pair[1] if pair[0] == leaf else pair[0]
        x = edges_list[up, hub]
        y = edges_list[hub, other]
        r = edges_list[tgt]
        del edges_list[up, hub]
        del edges_list[hub, other]
        del edges_list[tgt]
        u = random.random()
        edges_list[tgt[0], hub] = r * u
        edges_list[hub, tgt[1]] = r * (1.0 - u)
        edges_list[up, other] = x + y
        hastings_ratio = r / (x + y)
    new_postorder = postorder(adjlist2nodes_dict(edges_list), root_node)[::-1]
    return edges_list, new_postorder, hastings_ratio


def mvDualSlider(pi):
    """Redistribute the mass of two random frequencies (mcmc_gamma.pyx:187-198)."""
    i, j = random.sample(range(pi.shape[0]), 2)
    total = pi[i] + pi[j]
    x = total * random.random()
    pi[i], pi[j] = x, total - x
    return pi, 0.0


# ------------------------------------------------------------------------- start state
def rtree():
    """Random topology by repeated joining of the last two entries of a shuffled list
    (mcmc_gamma.pyx:305-317)."""
    pool = list(config.TAXA)
    random.shuffle(pool)
    while len(pool) > 1:
        last = str(pool.pop())
        second_last = str(pool.pop())
        pool.insert(0, "(" + second_last + "," + last + ")")
        random.shuffle(pool)
    pool.append(";")
    return "".join(pool)


def newick2bl(t):
    """Edge dict of a Newick string, internal nodes numbered downwards from
    n_leaves + #'(' in order of their opening bracket (mcmc_gamma.pyx:265-303).
    Edges are inserted when a leaf is read or a bracket closes; the outermost bracket is
    the root (= n_nodes) and yields no edge."""
    n_nodes = len(t.split(",")) + t.count("(")
    next_id = n_nodes
    edges = {}
    stack = []
    body = t.replace(";", "").replace(" ", "")
    i, n = 0, len(body)
    while i < n:
        ch = body[i]
        if ch == "(":
            stack.append(next_id)
            next_id -= 1
            i += 1
        elif ch == ",":
            i += 1
        else:
            start = i + 1 if ch == ")" else i
            j = start
            while j < n and body[j] not in "(),":
                j += 1
            token = body[start:j]
            name, _, length = token.partition(":")
            v = float(length) if length else 1.0
            if ch == ")":
                node = stack.pop()
                if stack:
                    edges[stack[-1], node] = v
            else:
                edges[stack[-1], name] = v
            i = j
    return edges, n_nodes


def init_tree():
    """Random start tree with tips renamed to 1..N and Exp(mean 0.1) branch lengths drawn in
    dict order (mcmc_gamma.pyx:244-260)."""
    edge_dict, n_nodes = newick2bl(rtree())
    for parent, child in list(edge_dict):
        if child in config.TAXA:
            del edge_dict[parent, child]
            edge_dict[parent, config.TAXA.index(child) + 1] = 1
    for key in edge_dict:
        edge_dict[key] = random.expovariate(1.0 / bl_exp_scale)
    return edge_dict, n_nodes


def init_alpha_rate():
    return random.expovariate(scaler_alpha)  # mcmc_gamma.pyx:262-263


def init_pi_er():
    """Start frequencies and exchangeabilities from NumPy's global generator
    (mcmc_gamma.pyx:319-332): JC is uniform, F81/GTR ~ Dirichlet(1); `er` is always drawn."""
    S = config.N_CHARS
    if config.MODEL == "JC":
        pi = np.repeat(1.0 / S, S)
    elif config.MODEL in ("F81", "GTR"):
        pi = np.random.dirichlet(np.repeat(1, S))
    else:
        raise UnboundLocalError("cannot access local variable 'pi'")  # what the reference does
    er = np.random.dirichlet(np.repeat(1, S * (S - 1) // 2))
    return pi, er

"""Traversal indexing of the rooted binary tree kept as an adjacency dict
``{(parent, child): branch length}`` -- integer artefacts that must match the reference
exactly.  Node ids: tips 1..N in file order, internal N+1..2N-1."""
from . import config


def adjlist2nodes_dict(edges_dict):
    """parent -> [children], children in dict insertion order (mcmc_gamma.pyx:220-232)."""
    kids = {}
    for parent, child in edges_dict:
        if parent in kids:
            kids[parent].append(child)
        else:
            kids[parent] = [child]
    return kids


def adjlist2reverse_nodes_dict(edges_dict):
    """child -> parent (mcmc_gamma.pyx:234-242)."""
    return {child: parent for parent, child in edges_dict}


def postorder(nodes_dict, node):
    """Edges in the reference's recursive order: both edges of `node`, then the subtree of
    the first child, then of the second (mcmc_gamma.pyx:200-218).  Callers reverse the list
    so children come before parents.  Iterative (no recursion limit on 1024-taxon trees)."""
    n_taxa = config.N_TAXA
    out = []
    stack = [node]
    while stack:
        nd = stack.pop()
        first, second = nodes_dict[nd]
        out.append((nd, first))
        out.append((nd, second))
        if second > n_taxa:
            stack.append(second)
        if first > n_taxa:
            stack.append(first)
    return out


def get_path2root(parent_of, internal_node, root):
    """Ancestors of `internal_node`, nearest first, ending with `root` (mcmc_gamma.pyx:26-38)."""
    path = []
    node = internal_node
    while True:
        node = parent_of[node]
        path.append(node)
        if node == root:
            return path


def adjlist2newickBL(edges_list, nodes_dict, node):
    """Newick string with branch lengths, children joined by ', ' (mcmc_gamma.pyx:549-571)."""
    n_taxa = config.N_TAXA
    parts = []
    for child in nodes_dict[node][:2]:
        if child > n_taxa:
            label = adjlist2newickBL(edges_list, nodes_dict, child)
        else:
            label = config.TAXA[child - 1]
        parts.append(label + ":" + str(edges_list[node, child]))
    return "(" + ", ".join(parts) + ")"

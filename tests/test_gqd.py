"""CPU: the open quartet-distance tool (cybayes_b200/gqd.py; the reference's gqd.py:1-22 shells out to a closed `qdist`
binary) against a brute-force quartet classifier working on bipartitions, and on hand-checked cases."""
import itertools
import random

import numpy as np

from cybayes_b200 import gqd


def _random_newick(names, rng):
    pool = list(names)
    rng.shuffle(pool)
    while len(pool) > 1:
        a = pool.pop(rng.randrange(len(pool)))
        b = pool.pop(rng.randrange(len(pool)))
        pool.append(f"({a}:{rng.random():.3f},{b}:{rng.random():.3f})")
    return pool[0] + ";"


def _clades(newick):
    """Leaf sets under every internal node, by string scanning (independent of gqd.parse_newick)."""
    out, stack = [], []
    tok = ""
    for ch in newick:
        if ch == "(":
            stack.append(set())
            tok = ""
        elif ch in ",);":
            name = tok.split(":")[0].strip()
            if name and not name.replace(".", "").isdigit():
                stack[-1].add(name)
            tok = ""
            if ch == ")":
                done = stack.pop()
                out.append(frozenset(done))
                if stack:
                    stack[-1] |= done
        else:
            tok += ch
    return out


def _brute_topology(clades, quartet):
    """ab|cd iff some clade separates exactly two of the four taxa from the other two."""
    q = set(quartet)
    for cl in clades:
        inside = q & cl
        if len(inside) == 2:
            a, b = sorted(inside)
            return frozenset([frozenset([a, b]), frozenset(q - inside)])
    return None


def test_quartet_topologies_match_brute_force():
    rng = random.Random(7)
    names = [f"t{i}" for i in range(11)]
    for _ in range(5):
        nw = _random_newick(names, rng)
        ch, nm, root = gqd.parse_newick(nw)
        D = gqd.lca_depths(ch, nm, root, tuple(names))
        quartets = np.array(list(itertools.combinations(range(len(names)), 4)), dtype=np.int32)
        top = gqd.quartet_topologies(D, quartets)
        clades = _clades(nw)
        for row, t in zip(quartets, top):
            a, b, c, d = (names[i] for i in row)
            want = _brute_topology(clades, (a, b, c, d))
            got = [frozenset([frozenset([a, b]), frozenset([c, d])]), frozenset([frozenset([a, c]), frozenset([b, d])]),
                   frozenset([frozenset([a, d]), frozenset([b, c])])][t] if t >= 0 else None
            assert got == want, (nw, row)


def test_gqd_hand_cases(tmp_path):
    gold = "((a:1,b:1):1,(c:1,d:1):1,e:1);"                      # unrooted-style trifurcation at the root
    same = "(((a,b),e),(c,d));"
    other = "((a,c),(b,d),e);"
    res = gqd.gqd(gold, [same, other])
    assert res[0] == 0.0
    # gold resolves ab|cd, ab|ce, ab|de, cd|ae, cd|be (5 butterflies); `other` says ac|bd, ac|.. etc.
    assert 0.0 < res[1] <= 1.0
    # a star gold tree resolves nothing
    assert gqd.gqd("(a,b,c,d,e);", [same])[0] == 0.0
    # names are cleaned like the reference does ('_' and '-' dropped), CLI prints mean std count over the second half
    trees = tmp_path / "run.trees"
    trees.write_text("".join(f"{i}\t{t}\n" for i, t in enumerate([other, other, "((a_,b-),(c,d),e);", same])))
    g = tmp_path / "gold.tre"
    g.write_text(gold)
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gqd.main([str(g), str(trees)])
    assert buf.getvalue().split() == ["0.0", "0.0", "2"]

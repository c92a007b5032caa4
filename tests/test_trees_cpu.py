"""Host-side tree reconstruction (phylo_b200/trees.py) from the integer tables, checked on the oracle's sweep."""
import re

import numpy as np

from oracle import vcsmc_oracle as O
from phylo_b200 import trees
from vcsmc_test_helpers import random_params, refs_from_oracle, synthetic_genome


def _tables(N=7, K=12, S=15):
    g = synthetic_genome(N, S, seed=2, gaps=0.05)
    p = random_params(N, False, seed=1)
    U = O.Uniforms.draw(N, K, seed=3)
    lam_l, lam_r, Q, pi = [t.detach() for t in O.model_from_params(p)]
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U)
    lref, rref = refs_from_oracle(res, N, K)
    out = {"left_ref": lref, "right_ref": rref, "left_branches": res.left_branches.numpy(),
           "right_branches": res.right_branches.numpy()}
    return res, out, N, K


def test_final_trees_contain_every_taxon_once_and_the_sampled_branches():
    res, out, N, K = _tables()
    taxa = ["t%d" % i for i in range(N)]
    for k in range(K):
        nwk = trees.final_tree_newick(k, taxa, out, fmt="%.17g")
        assert nwk.endswith(";") and nwk.count("(") == N - 1 and nwk.count(",") == N - 1
        assert sorted(re.findall(r"t\d+", nwk)) == sorted(taxa)
        root = N + (N - 2) * K + k
        assert trees.leaf_set(root, N, K, out["left_ref"], out["right_ref"]) == list(range(N))
        # every branch length in the string is one that was sampled at the event that created its parent
        merges = trees.merges_of(root, N, K, out["left_ref"], out["right_ref"], out["left_branches"], out["right_branches"])
        assert len(merges) == N - 1
        lengths = sorted(float(x) for x in re.findall(r":([0-9.eE+-]+)", nwk))
        expect = sorted([m[3] for m in merges] + [m[4] for m in merges])
        np.testing.assert_allclose(lengths, expect, rtol=1e-15)


def test_subtree_leaf_sets_match_the_oracles_leaf_counts():
    res, out, N, K = _tables()
    # the node created by slot k at the LAST rank event is the root of slot k's final forest: all N leaves
    assert all(int(c[-1]) == N for c in res.leaf_counts)
    for k in range(K):
        root = N + (N - 2) * K + k
        assert len(trees.leaf_set(root, N, K, out["left_ref"], out["right_ref"])) == N
    # an internal node's leaf set is the disjoint union of its children's
    for r in range(N - 1):
        for k in range(K):
            l, rr = int(out["left_ref"][r, k]), int(out["right_ref"][r, k])
            a = trees.leaf_set(l, N, K, out["left_ref"], out["right_ref"])
            b = trees.leaf_set(rr, N, K, out["left_ref"], out["right_ref"])
            assert not set(a) & set(b)
            assert trees.leaf_set(N + r * K + k, N, K, out["left_ref"], out["right_ref"]) == sorted(a + b)
    assert trees.newick(3, ["a", "b", "c", "d"], 1, None, None, None, None) == "d;"

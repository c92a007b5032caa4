"""Host-side tree reconstruction from the sweep's integer tables.

The reference carries taxa-label STRINGS through the TensorFlow graph (``vcsmc.py:306-313,:324,:424-425``: the
``jump_chain`` tensor) and, in ``csmc.py:175-215``, builds its trees from Python objects.  Here the device only keeps
integers -- for every rank event r and particle slot k the two children that were merged (``left_ref`` / ``right_ref``:
leaf i < N, or N + r'*K + k' for the node slot k' created at event r') and the two sampled branch lengths -- and the
trees are rebuilt on the host.  Node ids are global, so a subtree is followed through resampling without the ancestor
table.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def _children(node: int, n_taxa: int, n_particles: int, left_ref, right_ref, left_branches, right_branches):
    e = node - n_taxa
    r, k = divmod(e, n_particles)
    return int(left_ref[r, k]), int(right_ref[r, k]), float(left_branches[r, k]), float(right_branches[r, k])


def merges_of(node: int, n_taxa: int, n_particles: int, left_ref, right_ref, left_branches, right_branches
              ) -> List[Tuple[int, int, int, float, float]]:
    """Post-order list of (node, left, right, b_left, b_right) for the subtree rooted at ``node`` (iterative)."""
    out, stack = [], [(node, False)]
    while stack:
        x, done = stack.pop()
        if x < n_taxa:
            continue
        l, r, bl, br = _children(x, n_taxa, n_particles, left_ref, right_ref, left_branches, right_branches)
        if done:
            out.append((x, l, r, bl, br))
        else:
            stack.append((x, True))
            stack.append((r, False))
            stack.append((l, False))
    return out


def newick(node: int, taxa: Sequence[str], n_particles: int, left_ref, right_ref, left_branches, right_branches,
           fmt: str = "%.6g") -> str:
    """Newick string (with branch lengths) of the subtree rooted at ``node``."""
    n_taxa = len(taxa)
    if node < n_taxa:
        return str(taxa[node]) + ";"
    text: Dict[int, str] = {}
    for x, l, r, bl, br in merges_of(node, n_taxa, n_particles, left_ref, right_ref, left_branches, right_branches):
        ls = str(taxa[l]) if l < n_taxa else text.pop(l)
        rs = str(taxa[r]) if r < n_taxa else text.pop(r)
        text[x] = "(" + ls + ":" + (fmt % bl) + "," + rs + ":" + (fmt % br) + ")"
    return text[node] + ";"


def final_tree_newick(k: int, taxa: Sequence[str], out: Dict[str, np.ndarray], fmt: str = "%.6g") -> str:
    """The tree particle slot ``k`` holds after the last rank event, from ``VCSMC.outputs()`` / ``Sweep.output`` tables."""
    lref, rref = np.asarray(out["left_ref"]), np.asarray(out["right_ref"])
    n_events, n_particles = lref.shape
    n_taxa = n_events + 1
    root = n_taxa + (n_events - 1) * n_particles + int(k)
    return newick(root, taxa, n_particles, lref, rref, np.asarray(out["left_branches"]), np.asarray(out["right_branches"]), fmt)


def leaf_set(node: int, n_taxa: int, n_particles: int, left_ref, right_ref) -> List[int]:
    """Sorted leaf indices under ``node``."""
    leaves, stack = [], [node]
    while stack:
        x = stack.pop()
        if x < n_taxa:
            leaves.append(x)
        else:
            r, k = divmod(x - n_taxa, n_particles)
            stack.append(int(left_ref[r, k]))
            stack.append(int(right_ref[r, k]))
    return sorted(leaves)

"""Generate golden vectors by executing the REFERENCE's own code (csmc.py) in the build container.

Run once, here, where /root/reference exists:   python tests/golden/make_golden.py
Output (committed):  tests/golden/csmc_merge.npz, tests/golden/loader.npz

Nothing in the test-suite reads /root/reference at run time; the tests read the .npz.

What is executed live from the reference (matplotlib is stubbed because it is not installed,
csmc.py:15 imports it only for drawing):
  * CSMC.conditional_likelihood          csmc.py:300-309   (the merge formula)
  * CSMC.compute_log_conditional_likelihood  csmc.py:318-326   (full pruning on a Vertex tree)
  * CSMC.ncr                             csmc.py:155-160
The reference's runner.py loader is a __main__-local closure and cannot be imported, so
its golden is produced by exec'ing the exact source lines of form_dataset_from_strings
(runner.py:107-115) together with the alphabet dicts (runner.py:83-97).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_csmc():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import csmc  # noqa
    return csmc


def main():
    csmc = _import_csmc()
    rng = np.random.Generator(np.random.PCG64(20261018))
    S = 37
    taxa = ["t%d" % i for i in range(6)]
    codes = rng.integers(0, 4, (6, S))
    genome = np.eye(4)[codes]                       # one-hot [6,S,4]
    genome[2, 5] = 1.0                              # a gap column entry = all ones
    genome[4, 11] = 1.0
    model = csmc.CSMC({"taxa": taxa, "genome": genome})
    Q = model.Qmatrix.copy()
    prior = model.prior.copy()

    # ---- (1) merge formula on dense and one-hot children, several branch lengths
    L_l = rng.random((5, S, 4))
    L_r = rng.random((5, S, 4))
    L_l[0] = genome[0]; L_r[0] = genome[1]; L_l[1] = genome[2]
    b_l = np.array([2.0, 0.013, 0.4, 7.5, 1e-6])
    b_r = np.array([2.0, 0.9, 0.05, 0.31, 3.3])
    merged = np.zeros_like(L_l)
    for i in range(5):
        left, right = csmc.Vertex(data=L_l[i]), csmc.Vertex(data=L_r[i])
        merged[i] = model.conditional_likelihood(left, right, b_l[i], b_r[i])

    # ---- (2) full pruning log-likelihood on a fixed 6-taxon tree via Vertex objects
    # internal children go on the LEFT: csmc.py:273-275 tests curr.left.data_done before queueing curr.right
    merges = [(0, 1, 0.11, 0.23), (6, 2, 0.05, 0.4), (3, 4, 0.3, 0.02), (7, 8, 0.6, 0.17), (9, 5, 0.08, 0.9)]
    verts = {i: csmc.Vertex(id=taxa[i], data=genome[i]) for i in range(6)}
    for v in verts.values():
        v.data_done = True
    nid = 6
    for (l, r, bl, br) in merges:
        v = csmc.Vertex(id="n%d" % nid)
        v.left, v.right, v.left_branch, v.right_branch = verts[l], verts[r], bl, br
        verts[nid] = v
        nid += 1
    tree_loglik = model.compute_log_conditional_likelihood(verts[nid - 1])
    root_partials = verts[nid - 1].data

    ncr_vals = np.array([model.ncr(n, 2) for n in range(2, 13)])

    np.savez(os.path.join(HERE, "csmc_merge.npz"), Q=Q, prior=prior, genome=genome, L_l=L_l, L_r=L_r,
             b_l=b_l, b_r=b_r, merged=merged, merges=np.array(merges), tree_loglik=tree_loglik,
             root_partials=root_partials, ncr_vals=ncr_vals)

    # ---- (3) loader golden: exec the reference's own lines
    src = open(os.path.join(REF, "runner.py")).read().split("\n")
    ns = {"np": np}
    exec("\n".join(l[4:] for l in src[82:97]), ns)       # runner.py:83-97 alphabet dicts
    exec("\n".join(l[4:] for l in src[106:115]), ns)     # runner.py:107-115 form_dataset_from_strings
    strings = ["ACTTTGAGAG", "ACTTTGACAG", "ACT-TGACTG", "AC?TTGACTC"]  # runner.py:183 toy, with blanks
    dd = ns["form_dataset_from_strings"](strings, ns["Alphabet_dir_blank"])
    import pandas as pd
    prim = pd.read_pickle(os.path.join(REF, "data/primate.p"))
    pd_ = ns["form_dataset_from_strings"](list(prim.values()), ns["Alphabet_dir_blank"])
    np.savez_compressed(os.path.join(HERE, "loader.npz"), toy_strings=np.array(strings),
                        toy_genome=dd["genome"], toy_taxa=np.array(dd["taxa"]),
                        primate_strings=np.array(list(prim.values())),
                        primate_genome=pd_["genome"].astype(np.uint8))
    print("wrote golden vectors; tree_loglik =", tree_loglik)


if __name__ == "__main__":
    main()

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


# ------------------------------------------------------------------------------------------------------------
# Whole-sweep goldens: the reference's vcsmc.py / vncsmc.py executed UNMODIFIED under tests/golden/tf_shim.py
# ------------------------------------------------------------------------------------------------------------
def _ref_args(K, jc, M, nested):
    import types as _ty
    return _ty.SimpleNamespace(M=M, branch_prior=np.log(10), jcmodel=jc, optimizer="GradientDescentOptimizer",
                               dataset="golden", nested=nested, n_particles=K)


def _perturbed_variables(N, jc, rng, A=4):
    """Non-default variable values (named as vcsmc.py:119-124 names them) so that gradients are exercised away
    from the symmetric initial point."""
    ov = {"left_branches_param": np.log(10) + 0.3 * rng.standard_normal(N - 1),
          "right_branches_param": np.log(10) + 0.3 * rng.standard_normal(N - 1)}
    if not jc:
        ov["Qmatrix"] = 1.0 / A + 0.5 * rng.standard_normal((A, A))
        ov["Stationary_probs"] = 1.0 / A + 0.5 * rng.standard_normal(A)
    return ov


def run_reference_vcsmc(shim, mod, taxa, genome, K, jc, U, overrides):
    """Drives /root/reference/vcsmc.py: VCSMC.__init__ (:110-131) + sample_phylogenies (:406-451) with the
    randomness of `U` injected in the order body_rank_update (:332-400) consumes it."""
    import torch
    N = genome.shape[0]
    shim.reset()
    shim.variable_overrides.update(overrides)
    model = mod.VCSMC({"taxa": list(taxa), "genome": genome}, K, _ref_args(K, jc, 1, False))
    shim.feed(np.array([genome] * K, dtype=np.double))                      # vcsmc.py:479
    for r in range(N - 1):
        if r > 0:
            shim.push_uniforms("categorical", U.res[r][None, :])            # :285
        shim.push_uniforms("uniform", U.pair[r])                            # :303
        shim.push_uniforms("exponential", U.bl[r])                          # :355
        shim.push_uniforms("exponential", U.br[r])                          # :356
    model.sample_phylogenies()
    assert shim.uniforms_left() == 0
    grads = torch.autograd.grad(model.cost, shim.variables())
    out = _collect(shim, model, grads, N, K)
    tk = shim.trace["top_k"]                                                # two calls per rank event (:304-305)
    out["coal"] = np.concatenate([tk[2 * r].reshape(-1) for r in range(N - 1)])
    out["rem"] = np.concatenate([tk[2 * r + 1].reshape(-1) for r in range(N - 1)])
    return out


def run_reference_vncsmc(shim, mod, taxa, genome, K, M, jc, U, overrides):
    """Drives /root/reference/vncsmc.py: compute_potentials (:379-416), extend_partial_state (:295-322),
    body_rank_update (:432-499)."""
    import torch
    N = genome.shape[0]
    shim.reset()
    shim.variable_overrides.update(overrides)
    model = mod.VCSMC({"taxa": list(taxa), "genome": genome}, K, _ref_args(K, jc, M, True))
    shim.feed(np.array([genome] * K, dtype=np.double))
    for r in range(N - 1):
        if r > 0:
            shim.push_uniforms("categorical", U.res[r][None, :])            # vncsmc.py:285
        C = (N - r) * (N - r - 1) // 2
        for t in range(C):                                                  # :350-353, pair t in r1-major order
            shim.push_uniforms("exponential", U.look_bl[r][t])
            shim.push_uniforms("exponential", U.look_br[r][t])
        shim.push_uniforms("categorical", U.cat[r][:, None])                # :298
    model.sample_phylogenies()
    assert shim.uniforms_left() == 0
    grads = torch.autograd.grad(model.cost, shim.variables())
    out = _collect(shim, model, grads, N, K)
    cats = shim.trace["categorical"]
    out["choices"] = np.stack([c.reshape(-1) for c in cats if c.shape == (K, 1)])
    return out


def _collect(shim, model, grads, N, K):
    d = lambda t: t.detach().numpy().copy()
    out = {"elbo": float(model.elbo), "log_weights": d(model.log_weights), "log_likelihood": d(model.log_likelihood),
           "log_likelihood_tilde": d(model.log_likelihood_tilde), "log_likelihood_R": d(model.log_likelihood_R),
           "left_branches": d(model.left_branches), "right_branches": d(model.right_branches),
           "v_minus": d(model.v_minus).astype(np.int64), "Qmatrix": d(model.Qmatrix),
           "stationary_probs": d(model.stationary_probs),
           "jump_chains": np.array([[str(s) for s in row] for row in model.jump_chains])}
    anc = [c.reshape(-1) for c in shim.trace["categorical"] if c.shape == (1, K)]
    out["ancestors"] = np.stack([np.arange(K)] + anc).astype(np.int64)     # row 0 = identity (no resampling)
    for v, g in zip(shim.variables(), grads):
        out["var_" + v._shim_name] = d(v)
        out["dcost_" + v._shim_name] = d(g)                                 # d(cost) = -d(ELBO), vcsmc.py:447
    return out


SWEEP_CASES = [  # name, taxa subset, site slice, K, jc, perturbed variables, seed
    ("jc_n5_k16", 5, slice(0, 40), 16, True, False, 11),
    ("gtr_n5_k16", 5, slice(0, 40), 16, False, False, 12),
    ("jc_n8_k64_pert", 8, slice(100, 260), 64, True, True, 13),
    ("gtr_n8_k64_pert", 8, slice(100, 260), 64, False, True, 14),
    ("gtr_n12_k16_full", 12, slice(0, 898), 16, False, True, 15),
    ("jc_n12_k16_gaps", 12, slice(0, 898), 16, True, False, 16),
    ("gtr_n4_k2", 4, slice(0, 25), 2, False, True, 17),
    # a ONE-site slice cannot be run: the reference's own tf.squeeze at vcsmc.py:364-365 drops the site axis and
    # the concat at :368 raises (quirk Q9, confirmed under the shim); two sites is the smallest it accepts
    ("jc_n6_k32_twosites", 6, slice(7, 9), 32, True, True, 18),
    ("gtr_n7_k24_ties", 7, slice(300, 360), 24, False, True, 19),
]
NESTED_CASES = [  # name, taxa, sites, K, M, jc, perturbed, seed
    ("nested_jc_n5_k8_m2", 5, slice(0, 30), 8, 2, True, False, 21),
    ("nested_gtr_n5_k8_m5", 5, slice(0, 30), 8, 5, False, True, 22),
    ("nested_gtr_n7_k16_m2", 7, slice(200, 290), 16, 2, False, True, 23),
    ("nested_jc_n6_k12_m5", 6, slice(50, 110), 12, 5, True, True, 24),
]


def reference_sweeps():
    """tests/golden/ref_sweeps.npz: outputs of the reference's own sample_phylogenies + autodiff of its cost."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import tf_shim as shim
    shim.install()
    sys.path.insert(0, REF)
    import importlib
    ref_vcsmc = importlib.import_module("vcsmc")
    ref_vncsmc = importlib.import_module("vncsmc")
    assert ref_vcsmc.__file__.startswith(REF) and ref_vncsmc.__file__.startswith(REF)
    from oracle import vcsmc_oracle as O
    z = np.load(os.path.join(HERE, "loader.npz"))
    primate = z["primate_genome"].astype(np.float64)
    store = {}
    names = []
    for (name, n, sl, K, jc, pert, seed) in SWEEP_CASES:
        g = primate[:n, sl].copy()
        if name.endswith("gaps"):
            g[3, 10:20] = 1.0                                               # gap run = all-ones rows (runner.py:95)
        taxa = ["S%d" % i for i in range(n)]
        U = O.Uniforms.draw(n, K, seed=seed)
        if name.endswith("ties"):                                           # exact float32 ties inside rows
            for r in range(n - 1):
                U.pair[r][::3, 1] = U.pair[r][::3, 0]
                if U.pair[r].shape[1] > 3:
                    U.pair[r][1::4, 3] = U.pair[r][1::4, 2]
        ov = _perturbed_variables(n, jc, np.random.default_rng(seed)) if pert else {}
        out = run_reference_vcsmc(shim, ref_vcsmc, taxa, g, K, jc, U, ov)
        out.update(genome=g.astype(np.uint8), K=K, jc=int(jc), u_pair=np.concatenate([u.reshape(-1) for u in U.pair]),
                   u_bl=U.bl, u_br=U.br, u_res=U.res)
        print("%-22s ELBO %.9f" % (name, out["elbo"]))
        names.append(name)
        for k, v in out.items():
            store[name + "/" + k] = v
    nnames = []
    for (name, n, sl, K, M, jc, pert, seed) in NESTED_CASES:
        g = primate[:n, sl].copy()
        taxa = ["S%d" % i for i in range(n)]
        U = O.UniformsNested.draw(n, K, M, seed=seed)
        ov = _perturbed_variables(n, jc, np.random.default_rng(seed)) if pert else {}
        out = run_reference_vncsmc(shim, ref_vncsmc, taxa, g, K, M, jc, U, ov)
        out.update(genome=g.astype(np.uint8), K=K, M=M, jc=int(jc), u_cat=U.cat, u_res=U.res)
        for r in range(n - 1):
            out["u_look_bl_%d" % r] = U.look_bl[r]
            out["u_look_br_%d" % r] = U.look_br[r]
        print("%-22s ELBO %.9f" % (name, out["elbo"]))
        nnames.append(name)
        for k, v in out.items():
            store[name + "/" + k] = v
    store["cases"] = np.array(names)
    store["nested_cases"] = np.array(nnames)
    np.savez_compressed(os.path.join(HERE, "ref_sweeps.npz"), **store)
    print("wrote ref_sweeps.npz (%d + %d cases)" % (len(names), len(nnames)))


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "sweeps":
        main()
    reference_sweeps()

"""Shared helpers for the GPU parity tests (tests only; may import the oracle)."""
import numpy as np
import torch

from oracle import vcsmc_oracle as O


def gpu_uniforms(U: O.Uniforms, device="cuda"):
    """oracle.Uniforms -> the flat device arrays Sweep.set_uniforms takes."""
    u_pair = torch.from_numpy(np.concatenate([p.reshape(-1) for p in U.pair]).astype(np.float32)).to(device)
    return (u_pair, torch.from_numpy(U.bl).to(device).contiguous(), torch.from_numpy(U.br).to(device).contiguous(),
            torch.from_numpy(U.res).to(device).contiguous())


def random_params(N, jc, seed, scale=0.3):
    rng = np.random.default_rng(seed)
    p = O.Params.init(N, jc)
    p.left_branches_param = p.left_branches_param + torch.from_numpy(rng.normal(size=N - 1) * scale)
    p.right_branches_param = p.right_branches_param + torch.from_numpy(rng.normal(size=N - 1) * scale)
    if not jc:
        p.y_q = torch.from_numpy(rng.normal(size=(4, 4)) * scale)
        p.y_station = torch.from_numpy(rng.normal(size=4) * scale)
    return p


def synthetic_genome(N, S, seed=0, gaps=0.0):
    """i.i.d. uniform nucleotides (the reference's simulateDNA, runner.py:100-104), optional all-ones gap sites."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.eye(4)[rng.integers(0, 4, (N, S))]
    if gaps > 0:
        g[rng.random((N, S)) < gaps] = 1.0
    return g


def refs_from_oracle(res, N, K):
    """left/right child references of every event, from the oracle's coal positions and forest ids."""
    lref = np.zeros((N - 1, K), dtype=np.int64)
    rref = np.zeros((N - 1, K), dtype=np.int64)
    prev = np.tile(np.arange(N, dtype=np.int64), (K, 1))
    for r in range(N - 1):
        if r > 0:
            prev = res.forests[r - 1][res.ancestors[r]]
        lref[r] = prev[np.arange(K), res.coal[r][:, 0]]
        rref[r] = prev[np.arange(K), res.coal[r][:, 1]]
    return lref, rref


# ------------------------------------------------------------------------------------------------------------
# goldens produced by the reference's own vcsmc.py / vncsmc.py under tests/golden/tf_shim.py
# ------------------------------------------------------------------------------------------------------------
_REF_SWEEPS = None


def ref_sweeps():
    global _REF_SWEEPS
    if _REF_SWEEPS is None:
        import os
        _REF_SWEEPS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_sweeps.npz"))
    return _REF_SWEEPS


def ref_case_names(nested=False):
    return [str(s) for s in ref_sweeps()["nested_cases" if nested else "cases"]]


class RefCase:
    """One reference-produced sweep: inputs (genome, variables, uniforms) and every output the reference exposes."""

    def __init__(self, name):
        z = ref_sweeps()
        self.name = name
        self.z = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
        self.genome = self.z["genome"].astype(np.float64)
        self.N, self.S = self.genome.shape[:2]
        self.K = int(self.z["K"])
        self.jc = bool(int(self.z["jc"]))
        self.M = int(self.z["M"]) if "M" in self.z else None

    def __getitem__(self, k):
        return self.z[k]

    def params(self) -> O.Params:
        t = lambda k: torch.from_numpy(self.z[k].copy())
        if self.jc:
            return O.Params(t("var_left_branches_param"), t("var_right_branches_param"), None, None)
        return O.Params(t("var_left_branches_param"), t("var_right_branches_param"), t("var_Qmatrix"),
                        t("var_Stationary_probs"))

    def grads_elbo(self):
        """d(ELBO)/d(variables) in Params.tensors() order (the reference differentiates cost = -ELBO)."""
        names = ["left_branches_param", "right_branches_param"] + ([] if self.jc else ["Qmatrix", "Stationary_probs"])
        return [-self.z["dcost_" + n] for n in names]

    def uniforms(self):
        N, K = self.N, self.K
        if self.M is None:
            flat, pair, o = self.z["u_pair"], [], 0
            for r in range(N - 1):
                pair.append(flat[o:o + K * (N - r)].reshape(K, N - r).astype(np.float32))
                o += K * (N - r)
            return O.Uniforms(pair, self.z["u_bl"], self.z["u_br"], self.z["u_res"])
        return O.UniformsNested([self.z["u_look_bl_%d" % r] for r in range(N - 1)],
                                [self.z["u_look_br_%d" % r] for r in range(N - 1)], self.z["u_cat"], self.z["u_res"])

    def coal(self, r):
        """[K,2] pair positions of rank event r as the reference's tf.nn.top_k returned them (VCSMC only)."""
        N, K = self.N, self.K
        return self.z["coal"].reshape(N - 1, K, 2)[r]

    def rem(self, r):
        N, K = self.N, self.K
        o = sum(K * (N - q - 2) for q in range(r))
        return self.z["rem"][o:o + K * (N - r - 2)].reshape(K, N - r - 2)

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

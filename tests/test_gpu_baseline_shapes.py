"""Oracle parity (forward tables AND gradients) on the shapes BASELINE.json names, at oracle-sized K:
DS1 (27 taxa x 1949 sites, the real Hohna alignment the C3 shape comes from), 64 taxa (the C5 taxa count), and a
C2-shaped run (primate.p, general Q, K = 2048) on a site subset.  Lazy and eager schedules; plus the run-to-run
spread of the gradients, whose reverse merge accumulates with fp64 atomics."""
import os

import numpy as np
import pytest
import torch

from oracle import vcsmc_oracle as O
from vcsmc_test_helpers import gpu_uniforms, random_params, synthetic_genome
from test_gpu_sweep import compare_forward, compare_grads, oracle_param_grads, run_gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from phylo_b200 import ops as _ops
    return _ops


def oracle_case(name, jc):
    key = (name, jc)
    if key not in _cache:
        if name == "ds1":
            from phylo_b200.loader import load_dataset
            g = load_dataset("hohna_data_1", os.path.join(ROOT, "data"))["genome"]
            K = 32
        elif name == "n64":
            g = synthetic_genome(64, 200, seed=9, gaps=0.02)
            K = 32
        elif name == "c2":
            from phylo_b200.loader import load_dataset
            g = load_dataset("primate_data", os.path.join(ROOT, "data"))["genome"][:, 300:428]
            K = 2048
        N = g.shape[0]
        p = random_params(N, jc, seed=3) if name != "ds1" else O.Params.init(N, jc)
        U = O.Uniforms.draw(N, K, seed=17)
        res, g_ref = oracle_param_grads(g, K, p, U)
        _cache[key] = (g, K, p, U, res, g_ref)
    return _cache[key]


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("jc", [True, False])
@pytest.mark.parametrize("name", ["ds1", "n64", "c2"])
def test_baseline_shapes_match_oracle(ops, name, jc, lazy):
    g, K, p, U, res, g_ref = oracle_case(name, jc)
    N = g.shape[0]
    out, grads, _ = run_gpu(ops, g, K, p, U, jc, lazy=lazy)
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, jc)


def test_gradient_run_to_run_spread(ops):
    """The reverse merge accumulates child adjoints and dP with fp64 atomics, whose order varies between runs: the
    spread of the gradients over repeated backward passes of ONE forward stays at rounding level."""
    g, K, p, U, res, g_ref = oracle_case("c2", False)
    N, S = g.shape[0], g.shape[1]
    dev = lambda x: torch.as_tensor(x).cuda().contiguous()
    codes = ops.pack_alignment(dev(g))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    sw = ops.Sweep(N, S, K, False)
    sw.set_uniforms(*gpu_uniforms(U))
    sw.set_option("skip_zero", 0.0)   # the dense reverse sweep: every particle visited, the most atomics
    runs = []
    for it in range(6):
        sw.forward(codes, dev(lam_l), dev(lam_r), dev(Q), dev(pi.reshape(-1)))
        runs.append(np.concatenate([t.cpu().numpy().reshape(-1) for t in sw.backward(1.0)]))
    runs = np.stack(runs)
    scale = np.abs(runs).max(axis=0) + 1e-300
    spread = (runs.max(axis=0) - runs.min(axis=0)) / np.maximum(scale, 1e-3 * np.abs(runs).max())
    assert spread.max() <= 1e-12, spread.max()

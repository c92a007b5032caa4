"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vcsmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcsmc_[a-z0-9_]+)\s*\(", src)) - {"vcsmc_allreduce_fn", "vcsmc_comm_fn"})


def test_library_exports_every_declared_symbol():
    from phylo_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libvcsmc_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.vcsmc_abi_version() == 4


def test_arguments_are_validated_without_a_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    import ctypes as C
    from phylo_b200 import _lib
    lib = _lib.load()
    cfg = _lib.SweepConfig(1, 10, 4, 1, 1, 0)  # n_taxa = 1 is invalid
    sizes = _lib.SweepSizes()
    assert lib.vcsmc_sweep_query(C.byref(cfg), C.byref(sizes)) == _lib.ERR_ARG
    assert b"n_taxa" in lib.vcsmc_last_error()
    cfg = _lib.SweepConfig(12, 898, 2048, 0, 1, 0)
    assert lib.vcsmc_sweep_query(C.byref(cfg), C.byref(sizes)) == 0
    full = 11 * 2048 * 898 * 32
    assert sizes.retain_bytes > 2 * full and sizes.min_bytes < sizes.retain_bytes
    assert lib.vcsmc_merge_tiles(898) == 32 and lib.vcsmc_merge_tiles(10000) == 320   # upper bound of ell partials per particle
    with pytest.raises(_lib.VcsmcError):
        _lib.check(lib.vcsmc_transition_fwd(None, None, 4, 0, None, None))


def test_no_cpu_path():
    """CPU tensors are rejected: the product has no CPU fallback."""
    import torch
    from phylo_b200 import ops
    with pytest.raises(ValueError, match="no CPU path"):
        ops.pack_alignment(torch.zeros((2, 3, 4), dtype=torch.float64))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phylo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("# oracle", ""), f


def test_sweep_object_argument_checks_without_a_gpu():
    """Creating a sweep only carves a layout (no CUDA call that must succeed): the argument checks of the particle-
    sharding entry points and of the option table can be exercised on a CPU-only machine."""
    import ctypes as C
    from phylo_b200 import _lib
    lib = _lib.load()
    cfg = _lib.SweepConfig(8, 100, 64, 0, 1, 0, 0, 0)
    sizes = _lib.SweepSizes()
    assert lib.vcsmc_sweep_query(C.byref(cfg), C.byref(sizes)) == 0
    cfg.workspace_bytes = sizes.retain_bytes
    h = C.c_void_p()
    fake_ws = C.c_void_p(1 << 20)          # never dereferenced on the host
    assert lib.vcsmc_sweep_create(C.byref(cfg), fake_ws, C.byref(h)) == 0
    try:
        peers = (C.c_void_p * 8)(*([1 << 20] * 8))
        hook = _lib.COMM_FN(lambda *a: 0)
        assert lib.vcsmc_sweep_set_comm(h, 0, 9, hook, None, peers) == _lib.ERR_ARG          # more than 8 ranks
        assert lib.vcsmc_sweep_set_comm(h, 3, 3, hook, None, peers) == _lib.ERR_ARG          # rank out of range
        assert lib.vcsmc_sweep_set_comm(h, 0, 3, hook, None, peers) == _lib.ERR_ARG          # 64 particles / 3 ranks
        assert b"divisible" in lib.vcsmc_last_error()
        assert lib.vcsmc_sweep_set_comm(h, 0, 2, hook, None, None) == _lib.ERR_ARG           # no peer table
        assert lib.vcsmc_sweep_set_option(h, b"no_such_option", 1.0) == _lib.ERR_ARG
        for name in (b"skip_zero", b"skip_below", b"lazy", b"leaf_patterns", b"graph", b"site_begin", b"site_end", b"scalar_share"):
            assert lib.vcsmc_sweep_set_option(h, name, 1.0) == 0, name
        assert lib.vcsmc_sweep_output(h, b"no_such_table") is None
        assert lib.vcsmc_sweep_output(h, b"log_weights") is not None
        # backward before forward is a state error, not a crash
        assert lib.vcsmc_sweep_backward(h, 1.0, fake_ws, fake_ws, fake_ws, fake_ws, None) == _lib.ERR_STATE
    finally:
        lib.vcsmc_sweep_destroy(h)
    # the nested proposal cannot be particle-sharded
    cfg = _lib.SweepConfig(8, 100, 64, 0, 1, 0, 5, 0)
    assert lib.vcsmc_sweep_query(C.byref(cfg), C.byref(sizes)) == 0
    cfg.workspace_bytes = sizes.retain_bytes
    h = C.c_void_p()
    assert lib.vcsmc_sweep_create(C.byref(cfg), fake_ws, C.byref(h)) == 0
    try:
        peers = (C.c_void_p * 8)(*([1 << 20] * 8))
        assert lib.vcsmc_sweep_set_comm(h, 0, 2, _lib.COMM_FN(lambda *a: 0), None, peers) == _lib.ERR_STATE
    finally:
        lib.vcsmc_sweep_destroy(h)


def test_tabulated_expm_matches_scipy_on_the_host():
    """The kernels' exp(tQ) -- Q^k / k! tabulated once, a Horner scheme in t / 2^s per matrix, s squarings
    (common.cuh::m4_expm_tq; tf.linalg.expm at vcsmc.py:183-184) -- run on the HOST through vcsmc_transition_host:
    the reference's initial Q, row-softmax rate matrices as vcsmc.py:138-148 builds them, branch lengths from 1e-12
    to 60 (0 to 9 squarings)."""
    import ctypes as C
    import numpy as np
    from scipy.linalg import expm
    from phylo_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)

    def rate_matrix(logits):
        off = np.exp(logits - logits.max(axis=1, keepdims=True)) * (1 - np.eye(4))
        off /= off.sum(axis=1, keepdims=True)
        return off - np.eye(4)

    t = np.concatenate([[0.0, 1e-12, 1e-6, 0.05, 0.1, 0.24, 0.26, 1.0, 7.3, 60.0], rng.exponential(0.1, 40)])
    for Q in [rate_matrix(np.zeros((4, 4))), rate_matrix(rng.normal(size=(4, 4))), rate_matrix(3 * rng.normal(size=(4, 4)))]:
        Q = np.ascontiguousarray(Q)
        P = np.empty((t.size, 4, 4))
        rc = lib.vcsmc_transition_host(Q.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), t.size, P.ctypes.data_as(C.c_void_p))
        assert rc == 0
        ref = np.stack([expm(Q * ti) for ti in t])
        np.testing.assert_allclose(P, ref, rtol=1e-13, atol=1e-16)
        np.testing.assert_allclose(P.sum(axis=2), 1.0, rtol=0, atol=1e-13)   # rows of a transition matrix (t = 60: nine squarings)
    with pytest.raises(_lib.VcsmcError):
        _lib.check(lib.vcsmc_transition_host(None, None, 1, None))

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
    assert lib.vcsmc_abi_version() == 2


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

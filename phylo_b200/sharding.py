"""Host-side logic of site sharding, shared by VCSMC and by the CPU (gloo) tests.

Every rank holds all K particles for a contiguous slice of the (mini)batch's site list; the forest log-likelihood
sums are all-reduced once per rank event; the site-independent gradient terms are contributed by rank 0 only.
"""
from __future__ import annotations

import numpy as np


def local_sites(site_idx: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The slice of a batch's site list that ``rank`` of ``world`` processes (contiguous, sizes differ by <= 1)."""
    site_idx = np.asarray(site_idx, dtype=np.int32)
    if world <= 1:
        return site_idx
    return np.array_split(site_idx, world)[rank]


def scalar_share(rank: int, world: int) -> float:
    """Fraction of the site-independent gradient terms this rank contributes before the gradient all-reduce."""
    return 1.0 if (world <= 1 or rank == 0) else 0.0

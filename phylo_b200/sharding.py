"""Host-side logic of the two multi-GPU layouts, shared by VCSMC and by the CPU (gloo) tests.

Particle sharding (the default, north_star): rank g owns the logical particles [g K/G, (g+1) K/G) on ALL sites; per
rank event the step record is all-gathered, every rank derives identical ancestors, missing nodes are pulled from
their owners.  The reverse sweep is sharded by SITE on the gathered tables (``site_slice``); the site-independent
gradient terms are contributed by rank 0 only (``scalar_share``) and the parameter gradients are summed.

Site sharding (``sharding="sites"``): every rank holds all K particles for a contiguous slice of the (mini)batch's
site list; the forest log-likelihood sums are all-reduced once per rank event.
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


def particle_range(n_particles: int, rank: int, world: int):
    """[k0, k1) of the logical particles ``rank`` owns; K must be divisible by the world size."""
    if n_particles % world != 0:
        raise ValueError("n_particles=%d is not divisible by %d ranks" % (n_particles, world))
    kl = n_particles // world
    return rank * kl, (rank + 1) * kl


def owner_of(k, n_particles: int, world: int):
    """Rank that owns logical particle(s) ``k``."""
    return np.asarray(k) // (n_particles // world)


def site_slice(n_sites: int, rank: int, world: int):
    """[s0, s1) of the sites whose reverse sweep ``rank`` runs under particle sharding (contiguous, sizes differ by <= 1)."""
    base, extra = divmod(int(n_sites), int(world))   # the first `extra` ranks get one more site (np.array_split's rule)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def choose_sharding(requested, n_particles: int, world: int, nested: bool) -> str:
    """'particles' unless asked otherwise or impossible (nested proposal, K not divisible): then 'sites'."""
    if world <= 1:
        return "none"
    if requested in ("sites", "particles"):
        if requested == "particles" and (nested or n_particles % world != 0):
            raise ValueError("particle sharding needs the VCSMC proposal and n_particles divisible by the number of ranks")
        return requested
    return "sites" if (nested or n_particles % world != 0) else "particles"


def shared_seed(seed, dist=None) -> int:
    """The run's seed: rank 0's (explicit, or drawn when ``seed`` is None) on every rank.  Particle sharding needs every
    rank to derive identical ancestors, pairs and branch lengths from the same counter-based uniforms, site sharding
    identical draws for all particles, and both the same site minibatches; the reference seeds nothing, so an unseeded
    multi-process run would otherwise diverge silently."""
    s = int(seed if seed is not None else np.random.SeedSequence().entropy % (2 ** 63))
    if dist is not None:
        box = [s]
        dist.broadcast_object_list(box, src=0)
        s = int(box[0])
    return s

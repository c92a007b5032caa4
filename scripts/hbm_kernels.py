"""Kernel-level bench of the HBM-bound merge kernels on ALL-DISTINCT children (pool far larger than L2).

    python scripts/hbm_kernels.py [--particles 4096] [--sites 10000] [--model gtr|jc] [--reps 5]

In the SMC sweep's own workloads (ESS ~ 1) every particle of a rank event descends from one or two ancestors, so the
children of a launch are a handful of nodes that live in L2 and no kernel's HBM fraction can be read off the sweep.
Here every particle k merges its OWN two internal children (slots k and K + k) into its own node (slot 2K + k) through
the same C-ABI entry points (vcsmc_merge_fwd / vcsmc_merge_bwd): compulsory traffic per particle.site is
    merge_fwd : read 2 x 32 B, write 32 B                                  =  96 B
    merge_bwd : read 2 x 32 B (children) + 32 B (node adjoint),
                read-modify-write 2 x 32 B child adjoints (RED: 2 x 64 B)    = 224 B
Times are CUDA events on the launching stream after warm-up; fractions are of MEASURED_PEAKS.json's hbm_gbs.
bench.py imports `run` and puts the result under roofline.hbm_kernels; profiles/ holds the ncu --set full captures of
the same launches (dram__bytes per launch next to these algorithmic bytes).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BYTES = {"merge_fwd": 96.0, "merge_bwd": 224.0}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run(K=4096, S=10000, jc=False, reps=5, device="cuda"):
    from phylo_b200 import ops
    F64, I32 = torch.float64, torch.int32
    peak, src = hbm_peak()
    g = torch.Generator(device=device).manual_seed(0)
    pool = torch.empty((3 * K, S, 4), dtype=F64, device=device)
    pool[:2 * K].uniform_(0.05, 1.0, generator=g)
    ar = torch.arange(K, dtype=I32, device=device)
    lsrc, rsrc, dst = ar, ar + K, ar + 2 * K
    t = torch.empty(2 * K, dtype=F64, device=device).uniform_(0.02, 0.3, generator=g)
    eye = torch.eye(4, dtype=F64, device=device)
    Q = ((1 - eye) / 3 - eye).contiguous()
    P = ops.transition_fwd(None if jc else Q, t, jc).reshape(K, 32).contiguous()
    pi = torch.full((4,), 0.25, dtype=F64, device=device)
    out = {"particles": K, "sites": S, "model": "jc" if jc else "gtr", "pool_gb": pool.numel() * 8 / 1e9,
           "children": "all distinct (slot k, K + k -> 2K + k)", "peak_gbs": peak, "peak_source": src, "kernels": {}}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: ops.merge_fwd(None, pool, lsrc, rsrc, dst, P, pi, S, jc))
    gpool = torch.zeros((3 * K, S, 4), dtype=F64, device=device)
    gpool[2 * K:].uniform_(-1.0, 1.0, generator=g)
    coef = torch.empty(K, dtype=F64, device=device).uniform_(-1.0, 1.0, generator=g)
    dP = torch.zeros((K, 32), dtype=F64, device=device)
    dpi = torch.zeros(4, dtype=F64, device=device)
    ms_b = timed(lambda: ops.merge_bwd(None, pool, gpool, lsrc, rsrc, dst, P, pi, coef, S, jc, dP, dpi))
    for name, m in (("merge_fwd", ms), ("merge_bwd", ms_b)):
        alg = BYTES[name] * K * S
        out["kernels"][name] = {"ms": m, "algorithmic_bytes": alg, "bytes_per_merge": BYTES[name],
                                "achieved_gbs": alg / (m * 1e-3) / 1e9, "frac": alg / (m * 1e-3) / 1e9 / peak,
                                "merges_per_s": K * S / (m * 1e-3)}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--sites", type=int, default=10000)
    ap.add_argument("--model", default="gtr", choices=["gtr", "jc"])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(a.particles, a.sites, a.model == "jc", a.reps)))

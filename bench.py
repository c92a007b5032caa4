"""Benchmark of the VCSMC hot path: fwd+grad sweeps on a synthetic alignment, particle.site merges/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ELBO-grad step (one forward SMC sweep + the reverse sweep, what sess.run([optimizer, cost])
does at vcsmc.py:534) over the whole synthetic alignment.  A sweep performs K*S*(N-1) particle.site merges
forward and the same number backward; `value` = K*S*(N-1) / seconds per step (fwd+grad), whole job.

Workload (BASELINE.json configs[4], the one the metric's target is quoted on): 64 taxa x 10,000 sites x 65,536
particles, i.i.d. uniform nucleotides (numpy PCG64 seed 0), reference initial parameters (rates 10, `GTR' logits
1/4, uniform pi), float64.  N GPUs: the PARTICLES are sharded, K/N per GPU (strong scaling: K and S stay fixed; per
rank event one all-gather of the step record, nodes of remote ancestors pulled over NVLink, reverse sweep sharded by
site).  --sharding sites selects the site-sharded layout instead.

The default path is the production one: the forward scores every particle and materialises only the particles that the
next resampling draws ("lazy"); the reverse sweep skips events whose adjoint is exactly zero.  Results are identical to
the eager / dense schedule, which is also timed (N = 1) and reported under "eager_dense" with the HBM roofline
fractions of its two streaming kernels.

--impl reference times the restated reference (oracle/vcsmc_oracle.py: TensorFlow 1.15 cannot be installed
in this image) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "particle-site merges/sec (fwd+grad VCSMC sweep)"
UNIT = "merges/s"
BYTES_FWD, BYTES_BWD = 64.0, 128.0   # algorithmic bytes per merge, fp64 (SURVEY 8d / BASELINE.md section 3)
FP64_PEAK_TFMA = 18.2                # measured FP64 FMA rate of a B200 on this pool (scripts/microbench.cu), T op/s
# FP64-pipe operations per particle.site of the scoring path, by kind of merge: two internal children (16 DFMA + DADD +
# the mantissa DMUL), leaf + internal (4 DFMA + DADD + DMUL), two leaves (site patterns: O(1) per particle, counted as 0)
FP64_OPS_SCORE = {"gtr": (18.0, 6.0, 0.0), "jc": (6.0, 6.0, 0.0)}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--taxa", type=int, default=64)
    p.add_argument("--sites", type=int, default=10000)
    p.add_argument("--particles", type=int, default=65536)
    p.add_argument("--model", default="gtr", choices=["gtr", "jc"])
    p.add_argument("--dense", action="store_true", help="eager forward (every node stored) + reverse sweep without zero-adjoint skipping")
    p.add_argument("--sharding", default=None, choices=["particles", "sites"], help="multi-GPU layout (default: particles)")
    p.add_argument("--no-eager-dense", action="store_true", help="skip the extra eager/dense timing at N = 1")
    p.add_argument("--nested", type=int, default=0, help="M > 0: VNCSMC look-ahead proposal with M sub-samples (config 4)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-sample", default="64x1000x32", help="taxa x sites x particles of the CPU baseline sample")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arm: the restated reference (oracle) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample_run(sample: str, jc: bool, steps: int, warmup: int):
    """Time fwd+grad sweeps of the oracle on a bounded sample; returns (merges/s, cores, description, s/step)."""
    from oracle import vcsmc_oracle as O
    from phylo_b200.loader import synthetic_alignment
    n, s, k = [int(x) for x in sample.split("x")]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = synthetic_alignment(n, s)["genome"]
    p = O.Params.init(n, jc)
    times = []
    for it in range(warmup + steps):
        U = O.Uniforms.draw(n, k, seed=it)
        t0 = time.perf_counter()
        O.elbo_and_grads(g, k, p, U)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per = float(np.mean(times))
    desc = "%d taxa x %d sites x %d particles, %s, fwd+grad, restated reference (TensorFlow unavailable) in torch-CPU fp64" % (
        n, s, k, "JC" if jc else "GTR")
    return k * s * (n - 1) / per, cores, desc, per


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    jc = args.model == "jc"
    n, s, k = [int(x) for x in args.cpu_sample.split("x")]
    val, cores, desc, per = cpu_sample_run(args.cpu_sample, jc, max(args.steps, 1), args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "VCSMC %s fwd+grad sweep, 64 taxa x 10000 sites x 65536 particles (bounded sample: %s)" % (
            args.model.upper(), desc)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    from phylo_b200 import _lib, ops
    from phylo_b200.loader import synthetic_alignment
    from phylo_b200.sharding import site_slice
    from phylo_b200.vcsmc import VCSMC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO) to stdout by default: keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, S, K = args.taxa, args.sites, args.particles
    jc = args.model == "jc"

    class A:  # the argparse namespace the reference's class reads (vcsmc.py:111-120)
        M = max(args.nested, 1); branch_prior = float(np.log(10)); jcmodel = jc; optimizer = "GradientDescentOptimizer"
        dataset = "synthetic_%dx%d" % (N, S); nested = args.nested > 0; n_particles = K

    datadict = synthetic_alignment(N, S, seed=0)
    model = VCSMC(datadict, K, A, seed=0, sharding=args.sharding)
    sharding = model.sharding
    genome_host = torch.from_numpy(datadict["genome"]).pin_memory()
    variables = model.trainable_variables()

    split = []   # (forward ms, backward ms) of the steps run while `timing_split` is on

    def step(seed, timing_split=False):
        for v in variables:
            v.grad = None
        if timing_split:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        elbo = model.sample_phylogenies(need_grad=True, seed=seed)
        if timing_split:
            ev[1].record()
        (-elbo).backward()
        model._allreduce_grads()
        if timing_split:
            ev[2].record()
            torch.cuda.synchronize()
            split.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
        return elbo

    def step_e2e(seed):
        """Public-API step from HOST buffers: genome H2D + pack, parameters H2D, sweep fwd+grad, loss+grads D2H."""
        # (packed into the model's resident code buffer: same address every step, so the forward's CUDA graph is replayed)
        model.codes.copy_(ops.pack_alignment(genome_host.to(model.device, non_blocking=True)))
        host_params = [v.detach().cpu() for v in variables]
        for v, hp in zip(variables, host_params):
            v.data.copy_(hp.pin_memory(), non_blocking=True)
        elbo = step(seed)
        out = [elbo.detach().cpu()] + [v.grad.cpu() for v in variables]
        return out

    lazy_ok = not args.nested          # the nested proposal runs the eager forward

    def configure(dense):
        sw = model._last
        sw.set_option("skip_zero", 0.0 if dense else 1.0)
        if lazy_ok and world == 1:
            sw.set_option("lazy", 0.0 if dense else 1.0)

    step(1000)                          # creates the sweep engine
    sweep = model._last
    configure(args.dense)
    for w in range(max(args.warmup - 1, 2)):
        step(1001 + w)
    info = sweep.check_status()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM (the forward replays its captured CUDA graph)
    clocks = ClockSampler(local)
    launches0 = _lib.launch_count()
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        elbo = step(i)
    e1.record()
    barrier()
    clk = clocks.stop()
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    # ---- the same steps again with CUDA events around every merge launch (per-kernel times for the roofline; events
    #      cannot bracket kernels inside a graph, so this pass issues the launches one by one)
    sweep.set_option("profile", 1.0)
    barrier()
    for i in range(args.steps):
        step(i)
    barrier()
    prof = sweep.profile()
    sweep.set_option("profile", 0.0)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    merges = float(K) * S * (N - 1)
    if args.nested:   # every (pair, sub-sample) of the look-ahead is a particle-site merge too
        merges += float(K) * S * args.nested * sum((N - r) * (N - r - 1) // 2 for r in range(N - 1))
    value = merges / (ms_step * 1e-3)

    # ---- end-to-end region: host buffers in, loss + grads out, every step
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = merges / (float(te.item()) / args.steps)
    for i in range(2):   # forward / backward split of two more (graph-replayed) steps, device time
        step(i, timing_split=True)
    nparam = sum(v.numel() for v in variables)
    h2d = int(genome_host.numel() * 8 + nparam * 8)
    d2h = int(8 + nparam * 8)

    # ---- roofline of the dominant kernel (per-launch CUDA-event times recorded inside the timed region, this rank)
    S_fwd = len(model._local_sites(None))                        # sites this rank's forward covers
    K_fwd = K // world if sharding == "particles" else K         # particles this rank's forward covers
    if sharding == "particles":
        b0, b1 = site_slice(S, rank, world)
        S_bwd = b1 - b0
    else:
        S_bwd = S_fwd
    fwd_merges = float(K_fwd) * S_fwd * (N - 1) * args.steps
    bwd_merges = float(K) * S_bwd * (N - 1) * args.steps
    alg = {"merge_fwd": BYTES_FWD * fwd_merges, "merge_fwd_recompute": BYTES_FWD * bwd_merges,
           "merge_bwd": BYTES_BWD * bwd_merges, "materialise": BYTES_FWD * fwd_merges}
    dom = max(prof, key=lambda k: prof[k][0])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom_ms, dom_n = prof[dom]
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    lazy_fwd = lazy_ok and not (args.dense and world == 1)
    kname = "merge_score" if (dom == "merge_fwd" and lazy_fwd) else dom
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(kname + ("_jc" if jc else "_gtr"))
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "launches": dom_n,
                "avg_launch_ms": dom_ms / max(dom_n, 1),
                "algorithmic_bytes_per_launch": alg[dom] / max(dom_n, 1),
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "timing": "CUDA events around every launch of the kernel on the launching stream, in a repeat of the timed steps "
                          "(same seeds) issued launch by launch; the timed region itself replays the forward as a CUDA graph",
                "note": "achieved = algorithmic bytes (64 B/merge forward, 128 B/merge backward, fp64; SURVEY 8d) / kernel "
                        "time.  The lazy forward does not move those bytes: merge_score stores nothing and reads the shared "
                        "children from L2, so it is bounded by the FP64 pipe (see fp64), not by HBM; the HBM-bound schedule "
                        "is timed under eager_dense"}
    if kname == "merge_score" and dom_ms > 0:
        # what the scoring kernels really had to evaluate in the last sweep: classify its merges by kind of children
        lr, rr = sweep.output("left_ref"), sweep.output("right_ref")
        k_lo, k_hi = (rank * K_fwd, (rank + 1) * K_fwd) if sharding == "particles" else (0, K)
        l_leaf, r_leaf = (lr[:, k_lo:k_hi] < N), (rr[:, k_lo:k_hi] < N)
        n_ll = int((l_leaf & r_leaf).sum()); n_li = int((l_leaf ^ r_leaf).sum()); n_ii = int((~l_leaf & ~r_leaf).sum())
        o_ii, o_li, o_ll = FP64_OPS_SCORE[args.model]
        ops_per_sweep = float(S_fwd) * (o_ii * n_ii + o_li * n_li + o_ll * n_ll)
        tops = ops_per_sweep * args.steps / (dom_ms * 1e-3) / 1e12
        roofline["fp64"] = {"achieved": tops, "peak": FP64_PEAK_TFMA, "unit": "T FP64 op/s", "frac": tops / FP64_PEAK_TFMA,
                            "merges_by_children": {"internal+internal": n_ii, "leaf+internal": n_li, "leaf+leaf (site patterns)": n_ll},
                            "ops_per_merge": {"internal+internal": o_ii, "leaf+internal": o_li, "leaf+leaf": o_ll},
                            "peak_source": "scripts/microbench.cu on this pool's B200s (DFMA, 64 warps/SM)"}

    # ---- the eager / dense schedule (every node stored, no zero-adjoint skipping): the HBM-bound kernels
    eager = None
    if world == 1 and lazy_ok and not args.dense and not args.no_eager_dense:
        configure(True)
        step(2000)
        sweep.set_option("profile", 1.0)
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nd = 2
        d0.record()
        for i in range(nd):
            step(i)
        d1.record()
        barrier()
        pd = sweep.profile()
        sweep.set_option("profile", 0.0)
        dms = d0.elapsed_time(d1) / nd
        m1 = float(K) * S * (N - 1) * nd
        eager = {"ms_per_step": dms, "value": float(K) * S * (N - 1) / (dms * 1e-3), "unit": UNIT, "steps": nd,
                 "kernel_ms_per_step": {k: v[0] / nd for k, v in pd.items()},
                 "hbm_frac": {k: (b * m1 / (pd[k][0] * 1e-3) / 1e9 / peak if pd[k][0] > 0 else None)
                              for k, b in (("merge_fwd", BYTES_FWD), ("merge_bwd", BYTES_BWD))}}
        configure(False)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s %s fwd+grad sweep, %d taxa x %d sites x %d particles, i.i.d. uniform nucleotides "
                               "(PCG64 seed 0), reference initial parameters" % (
                                   "VNCSMC(M=%d)" % args.nested if args.nested else "VCSMC", args.model.upper(), N, S, K),
                   "taxa": N, "sites": S, "particles": K, "model": args.model, "sharding": "%s/%d" % (sharding, world),
                   "l2": "inputs larger than L2 (%.1f GB workspace; node pool and per-event tables streamed per sweep)" % (sweep.workspace.numel() / 1e9),
                   "forward": "lazy: every particle scored, survivors of the next resampling materialised (identical results)" if lazy_fwd else "eager: every node stored",
                   "backward": ("dense" if args.dense else "zero-adjoint events skipped (identical results)") + (", sharded by site" if sharding == "particles" else ""),
                   "nodes_retained": bool(sweep.retained), "backward_chunks": info["backward_chunks"],
                   "peak_pool_slots": info["peak_pool_slots"]},
        "roofline": roofline,
        "frac_of_fwdgrad_roofline": value * (BYTES_FWD + BYTES_BWD) / 1e9 / peak / world,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
        "elbo": float(elbo.detach()),
        "fwd_bwd_ms": [round(float(np.mean([a for a, _ in split])), 3), round(float(np.mean([b for _, b in split])), 3)],
    }
    if eager is not None:
        line["eager_dense"] = eager
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cval, cores, desc, per = cpu_sample_run(args.cpu_sample, jc, 1, 0)
        line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": desc + " (1 sweep, %.1f s)" % per}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        del sweep
        model._sweeps.clear()
        model._last = None
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)

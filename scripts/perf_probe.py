"""Quick device-side timing of one sweep shape (not the bench): python scripts/perf_probe.py N S K jc [dense|skip] [ws_gb|0] [M]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from phylo_b200 import ops
from phylo_b200.loader import synthetic_alignment

N, S, K, jc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
dense = len(sys.argv) > 5 and sys.argv[5] == "dense"
ws_gb = float(sys.argv[6]) if len(sys.argv) > 6 and float(sys.argv[6]) > 0 else None
M = int(sys.argv[7]) if len(sys.argv) > 7 else 0
g = synthetic_alignment(N, S)["genome"]
codes = ops.pack_alignment(torch.from_numpy(g).cuda())
lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
off = 1 - torch.eye(4, dtype=torch.float64, device="cuda")
Q = (off / 3 - torch.eye(4, dtype=torch.float64, device="cuda")).contiguous()
pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
t0 = time.time()
sw = ops.Sweep(N, S, K, jc, workspace_bytes=None if ws_gb is None else int(ws_gb * 2**30), n_sub=M)
if M:
    look = K * S * M * sum((N - r) * (N - r - 1) // 2 for r in range(N - 1))
    print("look-ahead merges per sweep: %.3e (+ %.3e main)" % (look, K * S * (N - 1)), flush=True)
print("workspace GB %.2f retained=%s min GB %.2f retain GB %.2f" % (sw.workspace.numel() / 2**30, sw.retained, sw.min_bytes / 2**30, sw.retain_bytes / 2**30), flush=True)
sw.set_seed(0)
sw.set_option("skip_zero", 0.0 if dense else 1.0)
sw.set_option("profile", 1.0)
sw.set_option("lazy", float(os.environ.get("LAZY", "1")))
merges = K * S * (N - 1)
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    elbo = sw.forward(codes, lam, lam, None if jc else Q, pi)
    e[1].record()
    grads = sw.backward(1.0)
    e[2].record()
    torch.cuda.synchronize()
    f, b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    prof = sw.profile()
    info = sw.check_status()
    print("it%d fwd %.2f ms bwd %.2f ms | fwd %.2f G/s  fwd+bwd %.2f G merges/s (%.1f%% of 33.5) | elbo %.3f | %s | %s" % (
        it, f, b, merges / f / 1e6, merges / (f + b) / 1e6, 100 * merges / (f + b) / 1e6 / 33.5, float(elbo), prof, info), flush=True)
ess = sw.output("ess").cpu().numpy()
print("ESS min/median/max", ess.min(), np.median(ess), ess.max())
print("dlam_l[:3]", grads[0][:3].cpu().numpy())

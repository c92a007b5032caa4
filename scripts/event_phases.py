"""Phase times of the lazy forward's event kernel: python scripts/event_phases.py N S K jc"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from phylo_b200 import ops
from phylo_b200.loader import synthetic_alignment
N, S, K, jc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
codes = ops.pack_alignment(torch.from_numpy(synthetic_alignment(N, S)["genome"]).cuda())
lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
eye = torch.eye(4, dtype=torch.float64, device="cuda")
Q = ((1 - eye) / 3 - eye).contiguous()
pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
sw = ops.Sweep(N, S, K, jc, keep_for_backward=False)
sw.set_seed(0)
sw.set_option("event_timing", 1.0)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sw.forward(codes, lam, lam, None if jc else Q, pi)
    e1.record()
    torch.cuda.synchronize()
    print("forward %.3f ms" % e0.elapsed_time(e1))
t = sw.event_timing().astype(np.float64)
names = ["weights", "unpack+max", "w+live", "-", "scan", "anc+rows", "survivors+alloc", "materialise", "propose", "offsets+sync", "scatter+pull"]
mid = t[1:N - 1]                       # full launches (r = 1 .. N-2)
d = np.diff(mid[:, :12], axis=1) / 1e3
print("per launch, us (median / mean over r = 1..N-2):")
for i, nm in enumerate(names):
    print("  %-18s %7.1f %7.1f" % (nm, np.median(d[:, i]), d[:, i].mean()))
print("  %-18s %7.1f %7.1f" % ("kernel total", np.median(mid[:, 11] - mid[:, 0]) / 1e3, (mid[:, 11] - mid[:, 0]).mean() / 1e3))
gap = (t[2:N - 1, 0] - t[1:N - 2, 11]) / 1e3   # end of launch r -> start of launch r+1: scoring + launch gaps
print("  %-18s %7.1f %7.1f" % ("between launches", np.median(gap), gap.mean()))

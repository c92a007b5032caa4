"""Phase times of the event kernel under particle sharding:
   python -m torch.distributed.run --nproc-per-node G scripts/event_phases_mp.py N S K jc"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from phylo_b200 import ops
from phylo_b200.comm import Comm
from phylo_b200.loader import synthetic_alignment
N, S, K, jc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
codes = ops.pack_alignment(torch.from_numpy(synthetic_alignment(N, S)["genome"]).cuda())
lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
eye = torch.eye(4, dtype=torch.float64, device="cuda")
Q = ((1 - eye) / 3 - eye).contiguous()
pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
comm = Comm()
sw = ops.Sweep(N, S, K, jc, keep_for_backward=False, comm=comm)
sw.set_seed(0)
sw.set_option("event_timing", 1.0)
for it in range(4):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sw.forward(codes, lam, lam, None if jc else Q, pi)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("forward %.3f ms" % e0.elapsed_time(e1), flush=True)
sw.check_status()
t = sw.event_timing().astype(np.float64)
names = ["weights+pack", "xsync1+unpack", "(max read)", "sumexp", "w+live", "scan+sync", "anc+rows", "survivors+alloc", "materialise", "propose", "offsets+xsync2", "scatter+pull"]
mid = t[1:N - 1]
d = np.diff(mid[:, :12], axis=1) / 1e3
out = ["rank %d per launch, us (median / mean over r = 1..N-2):" % rank]
lab = ["weights+pack", "xsync1+unpack", "w+live", "-", "scan", "anc+rows", "survivors+alloc", "materialise", "propose", "offsets+xsync2", "scatter+pull"]
for i, nm in enumerate(lab):
    out.append("  %-18s %7.1f %7.1f" % (nm, np.median(d[:, i]), d[:, i].mean()))
for nm, i0, i1 in (("  W + zeroing", 0, 12), ("  grid+signal+wait", 12, 13), ("  unpack", 13, 14), ("  grid sync", 14, 1)):
    dd = (mid[:, i1] - mid[:, i0]) / 1e3
    out.append("  %-18s %7.1f %7.1f" % (nm, np.median(dd), dd.mean()))
out.append("  %-18s %7.1f %7.1f" % ("kernel total", np.median(mid[:, 11] - mid[:, 0]) / 1e3, (mid[:, 11] - mid[:, 0]).mean() / 1e3))
gap = (t[2:N - 1, 0] - t[1:N - 2, 11]) / 1e3
out.append("  %-18s %7.1f %7.1f" % ("between launches", np.median(gap), gap.mean()))
for r in range(world):
    if r == rank:
        print("\n".join(out), flush=True)
    dist.barrier()
del sw
dist.destroy_process_group()

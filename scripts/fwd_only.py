"""Two forward-only sweeps (for an ncu launch list of one forward sweep): python scripts/fwd_only.py N S K jc [bwd]"""
import sys
import torch
sys.path.insert(0, ".")
from phylo_b200 import ops, _lib
from phylo_b200.loader import synthetic_alignment
N, S, K, jc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
bwd = len(sys.argv) > 5
codes = ops.pack_alignment(torch.from_numpy(synthetic_alignment(N, S)["genome"]).cuda())
lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
eye = torch.eye(4, dtype=torch.float64, device="cuda")
Q = ((1 - eye) / 3 - eye).contiguous()
pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
sw = ops.Sweep(N, S, K, jc, keep_for_backward=bwd)
sw.set_seed(0)
for it in range(2):
    n0 = _lib.launch_count()
    sw.forward(codes, lam, lam, None if jc else Q, pi)
    if bwd:
        sw.backward(1.0)
    torch.cuda.synchronize()
    print("launches per sweep (library count):", _lib.launch_count() - n0)

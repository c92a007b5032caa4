"""Consistency at scale: the general-Q kernels fed the JC rate matrix must reproduce the JC-specialised path."""
import sys
import torch
sys.path.insert(0, ".")
from phylo_b200 import ops
from phylo_b200.loader import synthetic_alignment

for (N, S, K) in [(12, 898, 2048), (27, 1949, 8192), (64, 1000, 4096), (64, 10000, 4096), (32, 10000, 65536), (64, 10000, 65536)]:
    g = synthetic_alignment(N, S)["genome"]
    codes = ops.pack_alignment(torch.from_numpy(g).cuda())
    lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
    eye = torch.eye(4, dtype=torch.float64, device="cuda")
    Q = (torch.full((4, 4), 0.25, dtype=torch.float64, device="cuda") - eye).contiguous()
    pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
    res = {}
    for jc in (True, False):
        sw = ops.Sweep(N, S, K, jc, keep_for_backward=True)
        sw.set_seed(3)
        e = float(sw.forward(codes, lam, lam, None if jc else Q, pi))
        g_ = sw.backward(1.0)
        res[jc] = (e, g_[0].cpu(), sw.output("log_z").cpu().clone(), sw.check_status())
        del sw
        torch.cuda.empty_cache()
    d = (res[True][2] - res[False][2]).abs()
    print(N, S, K, "elbo jc %.6f general %.6f | max |dlogz| %.3e at r=%d | dlam rel %.2e | %s" % (
        res[True][0], res[False][0], float(d.max()), int(d.argmax()),
        float((res[True][1] - res[False][1]).abs().max() / res[True][1].abs().max()), res[False][3]), flush=True)

import sys
import numpy as np, torch
sys.path.insert(0, ".")
from phylo_b200 import ops
from phylo_b200.loader import synthetic_alignment
N, S = 64, 10000
g = synthetic_alignment(N, S)["genome"]
codes = ops.pack_alignment(torch.from_numpy(g).cuda())
lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
eye = torch.eye(4, dtype=torch.float64, device="cuda")
Q = ((1 - eye) / 3 - eye).contiguous()
pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
for K in (1024, 8192, 65536):
    sw = ops.Sweep(N, S, K, False, keep_for_backward=False)
    sw.set_seed(4)
    e = float(sw.forward(codes, lam, lam, Q, pi))
    lz = sw.output("log_z").cpu().numpy()
    lb = sw.output("left_branches").cpu().numpy(); rb = sw.output("right_branches").cpu().numpy()
    lw = sw.output("log_weights").cpu().numpy()
    anc = sw.output("ancestors").cpu().numpy()
    print(K, "elbo", e, "logz[:4]", lz[:4], "min logz", lz.min(), "max logz", lz.max(), int(lz.argmax()))
    r = int(lz.argmax()); k = int(lw[r].argmax())
    print("   step", r, "best k", k, "lw", lw[r, k], "bl", lb[r, k], "br", rb[r, k], "max b overall", lb.max(), rb.max())
    # transition matrix for the winning branch lengths
    P = ops.transition_fwd(Q, torch.tensor([lb[r, k], rb[r, k], lb.max()], dtype=torch.float64, device="cuda"), False).cpu().numpy()
    print("   P rows sum", P.sum(axis=2), "P max", P.max(), "P min", P.min())
    del sw; torch.cuda.empty_cache()

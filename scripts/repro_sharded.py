"""Two (or more) ranks sharing ONE GPU run a seeded particle-sharded forward + backward (debugging aid):
python scripts/repro_sharded.py N S K [world]"""
import os, sys
import numpy as np, torch
import torch.multiprocessing as mp
sys.path.insert(0, ".")


def worker(rank, world, N, S, K):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = "29911"
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phylo_b200 import ops
    from phylo_b200.comm import Comm
    from phylo_b200.loader import synthetic_alignment
    from phylo_b200.sharding import site_slice, scalar_share
    g = synthetic_alignment(N, S)["genome"]
    codes = ops.pack_alignment(torch.from_numpy(g).cuda())
    lam = torch.full((N - 1,), 10.0, dtype=torch.float64, device="cuda")
    eye = torch.eye(4, dtype=torch.float64, device="cuda")
    Q = ((1 - eye) / 3 - eye).contiguous()
    pi = torch.full((4,), 0.25, dtype=torch.float64, device="cuda")
    sw = ops.Sweep(N, S, K, False, comm=Comm(), workspace_bytes=int(os.environ.get("WS_GB", "6")) << 30)
    s0, s1 = site_slice(S, rank, world)
    sw.set_option("site_begin", float(s0)); sw.set_option("site_end", float(s1)); sw.set_option("scalar_share", scalar_share(rank, world))
    sw.set_seed(0)
    for it in range(2):
        elbo = sw.forward(codes, lam, lam, Q, pi)
        sw.backward(1.0)
        torch.cuda.synchronize()
        print("rank", rank, "it", it, "elbo", float(elbo), sw.check_status(), flush=True)
    del sw
    dist.destroy_process_group()


if __name__ == "__main__":
    N, S, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    world = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    mp.spawn(worker, args=(world, N, S, K), nprocs=world, join=True)

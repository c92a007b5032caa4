"""The reference's command line (runner.py:12-58, :197-212) over the B200 implementation.

    python runner.py --dataset=primate_data --n_particles=16 --batch_size=256 --learning_rate=0.001 \
                     --num_epoch=100 --jcmodel=true [--nested=true | --twisting=true]

Same flags and defaults.  ``--twisting`` (README.md:27 / BASELINE.json) is accepted as an alias of ``--nested``
(runner.py:46-48 only defines --nested).  The reference at HEAD imports a missing module ``vcsmc_jet`` and
hard-codes ``ginkgo = True`` (runner.py:77,186-206); the working configuration -- ``import vcsmc`` with the
chosen dataset -- is what this runner means.  Extra flags: --seed, --data_dir, --no_save, --unknown_as_gap.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np


def _bool(x):
    return str(x).lower() == "true"


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Variational Combinatorial Sequential Monte Carlo")
    p.add_argument("--dataset", help="benchmark dataset to use.", default="primate_data")
    p.add_argument("--n_particles", type=int, help="number of SMC samples.", default=10)
    p.add_argument("--batch_size", type=int, help="number of sites on genome per batch.", default=256)
    p.add_argument("--learning_rate", type=float, help="Learning rate.", default=0.001)
    p.add_argument("--num_epoch", type=int, help="number of epoches to train.", default=100)
    p.add_argument("--optimizer", type=str, help="Optimizer for Training", default="GradientDescentOptimizer")
    p.add_argument("--branch_prior", type=float, help="Hyperparameter for branch length initialization.", default=np.log(10))
    p.add_argument("--M", type=int, help="number of subparticles to compute look-ahead particles", default=10)
    p.add_argument("--nested", default=False, type=_bool)
    p.add_argument("--twisting", default=None, type=_bool, help="alias of --nested")
    p.add_argument("--jcmodel", default=False, type=_bool)
    p.add_argument("--memory_optimization", help="Use memory optimization?", default="on")
    p.add_argument("--seed", type=int, default=None, help="seed of the counter-based generator (reference: unseeded)")
    p.add_argument("--data_dir", default="data")
    p.add_argument("--no_save", action="store_true")
    p.add_argument("--unknown_as_gap", default=True, type=_bool,
                   help="characters outside the alphabet (N in DS7, . n in DS10/11) become all-ones; false: KeyError like the reference")
    args = p.parse_args(argv)
    if args.twisting is not None:
        args.nested = args.twisting
    return args


def main(argv=None):
    args = parse_args(argv)
    import torch
    if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    from .loader import load_dataset
    datadict = load_dataset(args.dataset, args.data_dir, unknown_as_gap=args.unknown_as_gap)
    # runner.py:197-206: --nested=true imports vncsmc (same class name); here one class with args.nested
    from . import vcsmc
    model = vcsmc.VCSMC(datadict, K=args.n_particles, args=args, seed=args.seed)
    return model.train(epochs=args.num_epoch, batch_size=args.batch_size, learning_rate=args.learning_rate,
                       memory_optimization=args.memory_optimization, save=not args.no_save)


if __name__ == "__main__":
    main()
    sys.exit(0)

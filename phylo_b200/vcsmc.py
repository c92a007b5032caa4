"""``VCSMC(datadict, K, args).train(...)``: the reference's class interface (vcsmc.py:103-645) over the CUDA sweep.

Same constructor, same four trainable variables (vcsmc.py:119-124), same ``train`` protocol (site minibatches
drawn once, last slice never trained on -- quirk Q8 --, per-epoch full-data evaluation, ``results.p`` /
``run_parameters.txt`` with the reference's keys).  What differs, by design: no TensorFlow graph; the sweep and
its gradient are hand-written sm_100a kernels (libvcsmc_b200.so); the alignment is packed once into device
codes instead of being replicated K-fold on the host (vcsmc.py:479); randomness comes from a counter-based
generator keyed by (seed, rank event, particle) because the reference seeds nothing.

Multi-GPU: when torch.distributed is initialised the PARTICLES are sharded across ranks (each GPU holds K/G particles
on all sites; one all-gather of the step record per rank event; nodes of remote ancestors are pulled over NVLink;
the reverse sweep is sharded by site on the gathered tables).  ``sharding="sites"`` (or VCSMC_SHARDING=sites, and
always for the nested proposal) shards the sites of every (mini)batch instead: each rank holds all K particles for
its sites and the only per-event collective is an all-reduce of the K new log-likelihood sums.
"""
from __future__ import annotations

import math
import os
import pickle
import random
from datetime import datetime
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from .sharding import choose_sharding, local_sites, scalar_share, shared_seed, site_slice

F64 = torch.float64


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


class VCSMC:
    """
    VCSMC takes as input a dictionary (datadict) with two keys:
     taxa: a list of n strings denoting taxa
     genome: a 3 tensor [N,S,A] of genomes for the n taxa one hot encoded (ambiguous = all ones)
    """

    def __init__(self, datadict, K, args=None, device: Optional[str] = None, seed: Optional[int] = None,
                 sharding: Optional[str] = None):
        self.args = args
        self.taxa = list(datadict["taxa"])
        self.genome_NxSxA = np.asarray(datadict["genome"], dtype=np.float64)
        self.K = int(K)
        self.M = getattr(args, "M", 10)
        self.N = len(self.genome_NxSxA)
        self.S = len(self.genome_NxSxA[0])
        self.A = len(self.genome_NxSxA[0, 0])
        if self.A != 4:
            raise NotImplementedError("phylo_b200 kernels are specialised for a 4-letter alphabet (got A=%d)" % self.A)
        if not torch.cuda.is_available():
            raise RuntimeError("phylo_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device or ("cuda:%d" % torch.cuda.current_device()))
        self.jcmodel = bool(getattr(args, "jcmodel", False))
        # --nested=true selects the look-ahead proposal of vncsmc.py (same class name and signatures there)
        self.nested = bool(getattr(args, "nested", False))
        branch_prior = float(getattr(args, "branch_prior", math.log(10.0)))
        # the reference's variables (vcsmc.py:119-124); exp()/softmax parameterisations are applied in _model()
        dev = self.device
        self.left_branches_var = torch.full((self.N - 1,), branch_prior, dtype=F64, device=dev, requires_grad=True)
        self.right_branches_var = torch.full((self.N - 1,), branch_prior, dtype=F64, device=dev, requires_grad=True)
        if not self.jcmodel:
            self.y_q = torch.full((self.A, self.A), 1.0 / self.A, dtype=F64, device=dev, requires_grad=True)
            self.y_station = torch.full((self.A,), 1.0 / self.A, dtype=F64, device=dev, requires_grad=True)
        else:
            self.y_q = None
            self.y_station = None
        dist = _dist()
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
        self.seed = shared_seed(seed, dist)   # rank 0's seed (explicit or drawn) on every rank
        self._slice_rng = random.Random(self.seed) if (dist or seed is not None) else random
        self._step_counter = 0
        self._sweeps: Dict[tuple, ops.Sweep] = {}
        self._code_bufs: Dict[int, torch.Tensor] = {}
        self.sharding = choose_sharding(sharding or os.environ.get("VCSMC_SHARDING"), self.K, self.world, self.nested)
        self._comm = None
        if self.sharding == "particles":
            from .comm import Comm
            self._comm = Comm()
        # (a) pack the alignment ONCE into 4-bit device codes
        self.codes = ops.pack_alignment(torch.from_numpy(self.genome_NxSxA).to(dev))

    # -- parameters ---------------------------------------------------------------------------
    def trainable_variables(self) -> List[torch.Tensor]:
        v = [self.left_branches_var, self.right_branches_var]
        if not self.jcmodel:
            v += [self.y_q, self.y_station]
        return v

    def get_Q(self) -> torch.Tensor:
        """vcsmc.py:138-148 (general) / :126-129 (JC)."""
        eye = torch.eye(self.A, dtype=F64, device=self.device)
        if self.jcmodel:
            return torch.full((self.A, self.A), 1.0 / self.A, dtype=F64, device=self.device) - eye
        off = 1.0 - eye
        e = torch.exp(self.y_q * off) * off
        q = e / e.sum(dim=1, keepdim=True)
        return q - torch.diag(q.sum(dim=1))

    def get_stationary_probs(self) -> torch.Tensor:
        """vcsmc.py:133-136, shape [1,A]."""
        if self.jcmodel:
            return torch.full((1, self.A), 1.0 / self.A, dtype=F64, device=self.device)
        return torch.softmax(self.y_station, dim=0).unsqueeze(0)

    def _model(self):
        return (torch.exp(self.left_branches_var), torch.exp(self.right_branches_var), self.get_Q(),
                self.get_stationary_probs().reshape(-1))

    # -- the sweep ----------------------------------------------------------------------------
    def _local_sites(self, site_idx: Optional[np.ndarray]) -> np.ndarray:
        idx = np.arange(self.S, dtype=np.int32) if site_idx is None else np.asarray(site_idx, dtype=np.int32)
        if self.sharding != "sites":
            return idx                       # particle sharding: every rank sweeps all sites for its own particles
        return local_sites(idx, self.rank, self.world)

    def _sweep_for(self, n_sites: int, need_grad: bool) -> ops.Sweep:
        key = (n_sites, need_grad)
        if key not in self._sweeps:
            if need_grad and (n_sites, False) in self._sweeps:
                del self._sweeps[(n_sites, False)]
            sw = ops.Sweep(self.N, n_sites, self.K, self.jcmodel, keep_for_backward=need_grad, device=self.device,
                           n_sub=self.M if self.nested else 0, comm=self._comm)
            if self.sharding == "sites":
                import torch.distributed as dist
                sw.set_allreduce(lambda t: dist.all_reduce(t))
                sw.set_option("scalar_share", scalar_share(self.rank, self.world))
            elif self.sharding == "particles":
                s0, s1 = site_slice(n_sites, self.rank, self.world)   # the reverse sweep is sharded by site
                sw.set_option("site_begin", float(s0))
                sw.set_option("site_end", float(s1))
                sw.set_option("scalar_share", scalar_share(self.rank, self.world))
            self._sweeps[key] = sw
        return self._sweeps[key]

    def sample_phylogenies(self, site_idx: Optional[np.ndarray] = None, need_grad: bool = True,
                           seed: Optional[int] = None) -> torch.Tensor:
        """Main sampling routine (vcsmc.py:406-451): one sweep over the given sites; returns the ELBO.

        With ``need_grad`` the result is differentiable w.r.t. ``trainable_variables()`` (reverse sweep kernels).
        """
        n_batch = self.S if site_idx is None else len(site_idx)
        if self.sharding == "sites" and n_batch < self.world:
            # decided from the batch length alone, so EVERY rank raises before any collective (no rank is left waiting)
            raise ValueError("site sharding needs at least one site per rank: batch of %d sites on %d GPUs"
                             % (n_batch, self.world))
        local = self._local_sites(site_idx)
        if site_idx is None and self.sharding != "sites":
            codes = self.codes
        else:
            # gathered into a resident buffer per batch length: the sweep sees the same address every step, so its
            # captured launch graph stays valid across minibatches
            gathered = ops.gather_sites(self.codes, torch.from_numpy(local).to(self.device))
            buf = self._code_bufs.get(gathered.shape[1])
            if buf is None:
                buf = self._code_bufs[gathered.shape[1]] = torch.empty_like(gathered)
            buf.copy_(gathered)
            codes = buf
        sw = self._sweep_for(len(local), need_grad)
        if seed is None:
            self._step_counter += 1
            seed = (self.seed + 0x9E3779B97F4A7C15 * self._step_counter) % (2 ** 64)
        sw.set_seed(seed)
        lam_l, lam_r, Q, pi = self._model()
        self._last = sw
        if need_grad:
            elbo = ops.sweep_elbo(sw, codes, lam_l, lam_r, None if self.jcmodel else Q, pi)
        else:
            with torch.no_grad():
                elbo = sw.forward(codes, lam_l.contiguous(), lam_r.contiguous(), None if self.jcmodel else Q.contiguous(),
                                  pi.contiguous()).clone().reshape(())
        self.elbo = elbo
        self.cost = -elbo
        return elbo

    def release(self) -> None:
        """Drops the sweep engines (and with them their device workspaces and peer mappings)."""
        self._sweeps.clear()
        self._last = None
        self.elbo = None      # (its autograd node holds the sweep it came from)
        self.cost = None

    def _allreduce_grads(self):
        if self.world > 1:
            import torch.distributed as dist
            grads = [v.grad for v in self.trainable_variables() if v.grad is not None]
            if not grads:
                return
            flat = torch.cat([g.reshape(-1) for g in grads])       # one collective for the <= 2(N-1)+20 doubles
            if self._comm is not None:
                self._comm.all_reduce(flat)
            else:
                dist.all_reduce(flat)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def outputs(self) -> Dict[str, np.ndarray]:
        """The tensors the reference evaluates per epoch (vcsmc.py:538-551), as numpy arrays."""
        sw = self._last
        sw.check_status()
        names = ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R", "left_branches",
                 "right_branches", "v_minus", "ancestors", "left_ref", "right_ref")
        return {n: sw.output(n).cpu().numpy().copy() for n in names}

    def jump_chains(self, out: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
        """Taxa-label bookkeeping of vcsmc.py:306-313,:324,:424-425 rebuilt on the host from the integer tables.

        Returns the [K, 1 + N + (N-1) + ... + 2] string array the reference calls ``jump_chains`` (first column '').
        The labels are built correctly per particle; the reference's vcsmc.py reads particle 0's labels for every
        particle (quirk Q6), which is not reproduced.
        """
        out = out or self.outputs()
        sw = self._last
        rem = sw.rem_positions()
        N, K = self.N, self.K
        label = {i: self.taxa[i] for i in range(N)}
        forest = np.tile(np.arange(N, dtype=np.int64), (K, 1))
        cols = [np.full((K, 1), "", dtype=object)]
        for r in range(N - 1):
            if r > 0:
                forest = forest[out["ancestors"][r]]
            cols.append(np.vectorize(label.get, otypes=[object])(forest))
            for k in range(K):
                label[N + r * K + k] = label[int(out["left_ref"][r, k])] + "+" + label[int(out["right_ref"][r, k])]
            new_id = (N + r * K + np.arange(K, dtype=np.int64))[:, None]
            forest = np.concatenate([np.take_along_axis(forest, rem[r].astype(np.int64), axis=1), new_id], axis=1)
        self.final_trees = np.vectorize(label.get, otypes=[object])(forest[:, 0])
        return np.concatenate(cols, axis=1)

    def jump_chain_of(self, k: int, out: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
        """Row ``k`` of ``jump_chains()`` alone ([1, cols]), without building the labels of all K particles: what large
        runs store (K*N > 65,536), where the full string array would take minutes and gigabytes on the host."""
        out = out or self.outputs()
        sw = self._last
        N, K = self.N, self.K
        anc, lref, rref = out["ancestors"], out["left_ref"], out["right_ref"]
        after, label = {}, {i: self.taxa[i] for i in range(N)}

        def forest_before(r, j):                 # slot j's forest at event r, after resampling (vcsmc.py:288)
            return list(range(N)) if r == 0 else forest_after(r - 1, int(anc[r][j]))

        def forest_after(r, j):                  # ... after the merge of event r (vcsmc.py:313)
            if (r, j) not in after:
                f = forest_before(r, j)
                after[(r, j)] = [f[p] for p in sw.rem_row(r, j)] + [N + r * K + j]
            return after[(r, j)]

        def name(x):
            if x not in label:
                r, j = divmod(x - N, K)
                label[x] = name(int(lref[r, j])) + "+" + name(int(rref[r, j]))
            return label[x]

        row = [""]
        for r in range(N - 1):
            row += [name(x) for x in forest_before(r, k)]
        return np.array([row], dtype=object)

    def newick(self, k: Optional[int] = None, out: Optional[Dict[str, np.ndarray]] = None) -> str:
        """Newick string (with the sampled branch lengths) of the tree particle slot ``k`` holds after the last sweep;
        default: the particle with the largest ``log_likelihood_R``.  Rebuilt on the host from the integer tables
        (phylo_b200/trees.py); the reference only carries '+'-joined label strings (vcsmc.py:306-313)."""
        from . import trees
        out = out or self.outputs()
        if k is None:
            k = int(np.argmax(out["log_likelihood_R"]))
        return trees.final_tree_newick(k, self.taxa, out)

    # -- training driver (vcsmc.py:453-645) -----------------------------------------------------
    def batch_slices(self, n_sites: int, batch_size: int):
        """vcsmc.py:453-464: a fixed random partition of the sites, drawn once."""
        sites_list = list(range(n_sites))
        num_batches = n_sites // batch_size
        slices = []
        for _ in range(num_batches):
            sampled = self._slice_rng.sample(sites_list, batch_size)   # seeded (and identical on every rank) when a seed is known
            slices.append(sampled)
            sites_list = list(set(sites_list) - set(sampled))
        if len(sites_list) != 0:
            slices.append(sites_list)
        return slices

    def train(self, epochs=100, batch_size=128, learning_rate=0.001, memory_optimization="on", save=True,
              verbose=True):
        """Run the train op and evaluate variables, like vcsmc.py:466-645 (``memory_optimization`` is accepted and
        ignored: it toggles a TensorFlow Grappler pass, vcsmc.py:474-477)."""
        K = self.K
        self.lr = learning_rate
        say = print if (verbose and self.rank == 0) else (lambda *a, **k: None)
        slices = self.batch_slices(self.S, batch_size)
        say("================= Dataset shape: KxNxSxA =================")
        say((K, self.N, self.S, self.A))
        say("==========================================================")
        opt_name = getattr(self.args, "optimizer", "GradientDescentOptimizer")
        if opt_name == "Adam":
            self.optimizer = torch.optim.Adam(self.trainable_variables(), lr=self.lr, betas=(0.9, 0.999), eps=1e-8)
        else:
            self.optimizer = torch.optim.SGD(self.trainable_variables(), lr=self.lr)

        init_elbo = float(self.sample_phylogenies(need_grad=False))
        say("===================\nInitial evaluation of ELBO:", round(init_elbo, 3))
        say("Initial jump chain:")                                   # vcsmc.py:498-499 (row 0)
        say(self.jump_chain_of(0)[0])
        say("===================")
        save_dir = None
        if save and self.rank == 0:
            tm = str(datetime.now())
            root = "./results/" + str(getattr(self.args, "dataset", "dataset")) + "/" + str(getattr(self.args, "nested", False)) + \
                "/" + str(getattr(self.args, "n_particles", K)) + "/"
            save_dir = root + (tm[:10] + "-" + tm[11:13] + tm[14:16] + tm[17:19]) + "/"
            os.makedirs(save_dir, exist_ok=True)
            with open(save_dir + "run_parameters.txt", "w") as rp:
                rp.write("Initial evaluation of ELBO : " + str(init_elbo) + "\n")
                for k, v in (vars(self.args).items() if self.args is not None else []):
                    rp.write(str(k) + " : " + str(v) + "\n")
                rp.write(str(self.optimizer))

        say("Training begins --")
        elbos, Qmatrices, left_branches, right_branches, jump_chain_evolution = [], [], [], [], []
        log_weights, ll, ll_tilde, ll_R = [], [], [], []
        for i in range(epochs):
            bt = datetime.now()
            for j in range(len(slices) - 1):                       # quirk Q8: the last slice is never trained on
                self.optimizer.zero_grad(set_to_none=True)
                cost = -self.sample_phylogenies(np.asarray(slices[j], dtype=np.int32), need_grad=True)
                cost.backward()
                self._allreduce_grads()
                self.optimizer.step()
            cost = -float(self.sample_phylogenies(need_grad=False))  # per-epoch full-data evaluation, vcsmc.py:538-551
            out = self.outputs()
            stats = self.get_stationary_probs().detach().cpu().numpy()
            Qs = self.get_Q().detach().cpu().numpy()
            lb_param = torch.exp(self.left_branches_var).detach().cpu().numpy()
            rb_param = torch.exp(self.right_branches_var).detach().cpu().numpy()
            # the reference stores all K rows (vcsmc.py:550,:589); large runs keep the best particle's row only
            jc = self.jump_chains(out) if K * self.N <= 1 << 16 else self.jump_chain_of(int(np.argmax(out["log_likelihood_R"])), out)
            say("Epoch", i + 1)
            say("ELBO\n", round(-cost, 3))
            say("Stationary probabilities\n", stats)
            say("Q-matrix\n", Qs)
            say("LB param:\n", lb_param)
            say("RB param:\n", rb_param)
            elbos.append(-cost)
            Qmatrices.append(Qs)
            left_branches.append(out["left_branches"])
            right_branches.append(out["right_branches"])
            ll.append(out["log_likelihood"])
            ll_tilde.append(out["log_likelihood_tilde"])
            ll_R.append(out["log_likelihood_R"])
            log_weights.append(out["log_weights"])
            jump_chain_evolution.append(jc)
            say("Time spent\n", datetime.now() - bt, "\n-----------------------------------------")
        say("Done training.")

        best = int(np.argmax(elbos)) if elbos else 0
        resultDict = {"cost": np.asarray(elbos), "nParticles": self.K, "nTaxa": self.N, "lr": self.lr,
                      "log_weights": np.asarray(log_weights), "Qmatrices": np.asarray(Qmatrices),
                      "left_branches": left_branches, "right_branches": right_branches, "log_lik": np.asarray(ll),
                      "ll_tilde": np.asarray(ll_tilde), "log_lik_R": np.asarray(ll_R),
                      "jump_chain_evolution": jump_chain_evolution, "best_epoch": best,
                      "best_log_lik": np.asarray(ll_R)[best] if ll_R else None,
                      "best_jump_chain": jump_chain_evolution[best] if jump_chain_evolution else None}
        if save_dir is not None:
            with open(save_dir + "results.p", "wb") as f:
                pickle.dump(resultDict, f)
        say("Finished...")
        self.save_dir = save_dir
        return resultDict

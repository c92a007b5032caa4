"""PyTorch-facing operators over the C ABI (include/vcsmc_b200.h).

torch is plumbing here: it owns device memory and streams.  Every function below forwards to a
hand-written sm_100a kernel in libvcsmc_b200.so through ctypes; nothing is computed by torch ops and
there is no CPU path (CPU tensors are rejected).

Custom ops registered under ``torch.ops.vcsmc``: ``transition``, ``merge``, ``propose_pairs``,
``resample`` -- ``transition`` and ``merge`` carry autograd (reverse pruning).  The whole SMC sweep with
its hand-written reverse sweep is ``sweep_elbo`` (a torch.autograd.Function over a ``Sweep`` object).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check

F64, F32, I32, U8 = torch.float64, torch.float32, torch.int32, torch.uint8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor: phylo_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


# ------------------------------------------------------------------------------------------
# (a) loader
# ------------------------------------------------------------------------------------------
def pack_alignment(genome: torch.Tensor) -> torch.Tensor:
    """[N,S,4] float64 0/1 state masks (runner.py:107-115) -> [N,S] uint8 4-bit codes, on device."""
    _chk(genome, F64, "genome")
    if genome.dim() != 3 or genome.shape[2] != 4:
        raise ValueError("genome must be [N,S,4] (alphabet size 4); got %s" % (tuple(genome.shape),))
    N, S = genome.shape[0], genome.shape[1]
    codes = torch.empty((N, S), dtype=U8, device=genome.device)
    status = torch.zeros(1, dtype=I32, device=genome.device)
    check(_lib.load().vcsmc_pack_alignment(_ptr(genome), N, S, _ptr(codes), _ptr(status), _stream()))
    rc = int(status.item())
    if rc != 0:
        raise _lib.VcsmcError(rc, "genome entries must be exactly 0/1 with at least one 1 per site")
    return codes


def gather_sites(codes: torch.Tensor, site_idx: torch.Tensor) -> torch.Tensor:
    """np.take(data, slice, axis=2) of vcsmc.py:533 on packed codes."""
    _chk(codes, U8, "codes"); _chk(site_idx, I32, "site_idx")
    N, S = codes.shape
    out = torch.empty((N, site_idx.numel()), dtype=U8, device=codes.device)
    check(_lib.load().vcsmc_gather_sites(_ptr(codes), N, S, _ptr(site_idx), site_idx.numel(), _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------
# (b) transition matrices
# ------------------------------------------------------------------------------------------
def transition_fwd(Q: Optional[torch.Tensor], t: torch.Tensor, jc: bool) -> torch.Tensor:
    _chk(t, F64, "t")
    if not jc:
        _chk(Q, F64, "Q")
    P = torch.empty((t.numel(), 4, 4), dtype=F64, device=t.device)
    check(_lib.load().vcsmc_transition_fwd(_ptr(Q), _ptr(t), t.numel(), int(jc), _ptr(P), _stream()))
    return P


def transition_bwd(Q: Optional[torch.Tensor], t: torch.Tensor, dP: torch.Tensor, jc: bool):
    """Returns (dt[n], dQ_each[n,4,4] or None).  In JC mode dP is the compressed adjoint (see the header)."""
    _chk(t, F64, "t"); _chk(dP, F64, "dP")
    n = t.numel()
    dt = torch.empty(n, dtype=F64, device=t.device)
    dQ = None if jc else torch.empty((n, 4, 4), dtype=F64, device=t.device)
    check(_lib.load().vcsmc_transition_bwd(_ptr(Q), _ptr(t), _ptr(dP), n, int(jc), _ptr(dt), _ptr(dQ), _stream()))
    return dt, dQ


# ------------------------------------------------------------------------------------------
# (c) merge
# ------------------------------------------------------------------------------------------
def merge_tiles(n_sites: int) -> int:
    return int(_lib.load().vcsmc_merge_tiles(n_sites))


def merge_fwd(codes: Optional[torch.Tensor], pool: torch.Tensor, lsrc: torch.Tensor, rsrc: torch.Tensor,
              dst: Optional[torch.Tensor], P: torch.Tensor, pi: torch.Tensor, n_sites: int, jc: bool) -> torch.Tensor:
    """Writes the new nodes into ``pool`` [slots, slot_sites, 4] and returns ell[K]."""
    _chk(pool, F64, "pool"); _chk(lsrc, I32, "lsrc"); _chk(rsrc, I32, "rsrc"); _chk(P, F64, "P"); _chk(pi, F64, "pi")
    K = lsrc.numel()
    ell_part = torch.empty(K * max(merge_tiles(n_sites), 1), dtype=F64, device=pool.device)
    ell = torch.empty(K, dtype=F64, device=pool.device)
    stride = codes.shape[1] if codes is not None else 0
    check(_lib.load().vcsmc_merge_fwd(_ptr(codes), stride, _ptr(pool), pool.shape[1], _ptr(lsrc), _ptr(rsrc), _ptr(dst),
                                      _ptr(P), _ptr(pi), K, n_sites, int(jc), _ptr(ell_part), _ptr(ell), _stream()))
    return ell


def merge_bwd(codes: Optional[torch.Tensor], pool: torch.Tensor, gpool: torch.Tensor, lsrc, rsrc, gsrc, P, pi, coef,
              n_sites: int, jc: bool, dP: torch.Tensor, dpi: Optional[torch.Tensor]) -> None:
    """Accumulates into gpool, dP [K,32], dpi [4] (summed over particles)."""
    K = lsrc.numel()
    stride = codes.shape[1] if codes is not None else 0
    check(_lib.load().vcsmc_merge_bwd(_ptr(codes), stride, _ptr(pool), _ptr(gpool), pool.shape[1], _ptr(lsrc), _ptr(rsrc),
                                      _ptr(gsrc), _ptr(P), _ptr(pi), _ptr(coef), K, n_sites, int(jc), _ptr(dP),
                                      _ptr(dpi), _stream()))


# ------------------------------------------------------------------------------------------
# (d) proposal / resampling / uniforms
# ------------------------------------------------------------------------------------------
def propose_pairs(u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _chk(u, F32, "u")
    K, n = u.shape
    coal = torch.empty((K, 2), dtype=I32, device=u.device)
    rem = torch.empty((K, max(n - 2, 0)), dtype=I32, device=u.device)
    check(_lib.load().vcsmc_propose_pairs(_ptr(u), K, n, _ptr(coal), _ptr(rem) if n > 2 else None, _stream()))
    return coal, rem


def resample(lw: torch.Tensor, u: torch.Tensor):
    """Returns (idx int32 [K], logsumexp(lw), ESS)."""
    _chk(lw, F64, "lw"); _chk(u, F64, "u")
    K = lw.numel()
    idx = torch.empty(K, dtype=I32, device=lw.device)
    out = torch.empty(2, dtype=F64, device=lw.device)
    work = torch.empty(int(_lib.load().vcsmc_resample_work_doubles(K)), dtype=F64, device=lw.device)
    check(_lib.load().vcsmc_resample(_ptr(lw), _ptr(u), K, _ptr(idx), out.data_ptr(), out.data_ptr() + 8, _ptr(work), _stream()))
    return idx, out[0], out[1]


def philox_step_uniforms(seed: int, r: int, k0: int, K: int, n: int, device="cuda"):
    """The uniforms rank event r consumes for logical particles k0..k0+K-1: (u_pair[K,n] f32, u_bl, u_br, u_res [K] f64)."""
    u_pair = torch.empty((K, n), dtype=F32, device=device)
    u_bl = torch.empty(K, dtype=F64, device=device)
    u_br = torch.empty(K, dtype=F64, device=device)
    u_res = torch.empty(K, dtype=F64, device=device)
    check(_lib.load().vcsmc_philox_step_uniforms(seed, r, k0, K, n, _ptr(u_pair), _ptr(u_bl), _ptr(u_br), _ptr(u_res), _stream()))
    return u_pair, u_bl, u_br, u_res


# ------------------------------------------------------------------------------------------
# the sweep
# ------------------------------------------------------------------------------------------
_F64_OUT = {"elbo": (1,), "log_weights": None, "log_likelihood": None, "log_likelihood_tilde": "K",
            "log_likelihood_R": "K", "left_branches": None, "right_branches": None, "log_z": "N1", "ess": "N1",
            "node_coef": None}
_I32_OUT = {"v_minus": "K", "ancestors": None, "left_ref": None, "right_ref": None, "leaf_counts": None, "status": (8,),
            "choice": None}


class Sweep:
    """One VCSMC sweep engine for a fixed (N, S, K, model) shape; owns a device workspace.

    forward(): sample_phylogenies (vcsmc.py:406-451).  backward(): the reverse sweep of
    ``optimizer.minimize(cost)`` (vcsmc.py:488-491).  ``workspace_bytes=None`` retains every node when
    that fits in ``mem_fraction`` of free device memory, otherwise uses the garbage-collected forward
    pool + site-chunked backward.
    """

    def __init__(self, n_taxa: int, n_sites: int, n_particles: int, jc: bool, keep_for_backward: bool = True,
                 workspace_bytes: Optional[int] = None, device="cuda", mem_fraction: float = 0.85, n_sub: int = 0,
                 comm=None):
        lib = _lib.load()
        self.N, self.S, self.K, self.jc, self.keep = int(n_taxa), int(n_sites), int(n_particles), bool(jc), bool(keep_for_backward)
        self.device = torch.device(device)
        self.M = int(n_sub)   # 0: VCSMC (vcsmc.py); M > 0: VNCSMC look-ahead with M sub-samples (vncsmc.py)
        cfg = _lib.SweepConfig(self.N, self.S, self.K, int(self.jc), int(self.keep), 0, self.M, 0)
        sizes = _lib.SweepSizes()
        check(lib.vcsmc_sweep_query(C.byref(cfg), C.byref(sizes)))
        self.min_bytes, self.retain_bytes = int(sizes.min_bytes), int(sizes.retain_bytes)
        if workspace_bytes is None:
            free, _total = torch.cuda.mem_get_info(self.device)
            budget = int(free * mem_fraction)
            if comm is not None and comm.world > 1:
                # particle sharding: every rank carves the same layout, and nodes are never all retained
                budget = comm.min_int(budget)
                workspace_bytes = max(min(budget, self.retain_bytes + (64 << 20)), self.min_bytes)
            elif self.retain_bytes <= budget:
                workspace_bytes = self.retain_bytes
            elif not self.keep and self.M == 0:
                # an evaluation-only lazy sweep stores the survivors' nodes and nothing else: the garbage-collected pool
                # of min_bytes (2 K slots) is all it can ever use -- leave the rest of the device to the training sweep
                workspace_bytes = self.min_bytes
            else:
                workspace_bytes = max(budget, self.min_bytes)
        workspace_bytes = (int(workspace_bytes) + 255) // 256 * 256
        self.workspace = torch.empty(workspace_bytes, dtype=U8, device=self.device)
        cfg.workspace_bytes = workspace_bytes
        h = C.c_void_p()
        check(lib.vcsmc_sweep_create(C.byref(cfg), self.workspace.data_ptr(), C.byref(h)))
        self._h = h
        self._lib = lib
        self._keepalive = None
        self._hook = None
        self._comm_hook = None
        self._peers = None
        self.retained = workspace_bytes >= self.retain_bytes
        if comm is not None and comm.world > 1:
            self.set_comm(comm)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.vcsmc_sweep_destroy(h)
        peers = getattr(self, "_peers", None)
        if peers is not None:
            peers.close()

    # -- configuration
    def set_seed(self, seed: int) -> None:
        check(self._lib.vcsmc_sweep_set_seed(self._h, int(seed)))

    def set_uniforms(self, u_pair: torch.Tensor, u_bl: torch.Tensor, u_br: torch.Tensor, u_res: torch.Tensor) -> None:
        """u_pair: ragged concat over r of [K, N-r] float32 (flat); u_bl/u_br/u_res: [N-1,K] float64."""
        _chk(u_pair, F32, "u_pair"); _chk(u_bl, F64, "u_bl"); _chk(u_br, F64, "u_br"); _chk(u_res, F64, "u_res")
        need = self.K * sum(self.N - r for r in range(self.N - 1))
        if u_pair.numel() != need or u_bl.numel() != (self.N - 1) * self.K:
            raise ValueError("uniform arrays have the wrong size")
        self._uniforms = (u_pair, u_bl, u_br, u_res)
        check(self._lib.vcsmc_sweep_set_uniforms(self._h, _ptr(u_pair), _ptr(u_bl), _ptr(u_br), _ptr(u_res)))

    def set_uniforms_nested(self, look_bl: torch.Tensor, look_br: torch.Tensor, cat: torch.Tensor, res: torch.Tensor) -> None:
        """VNCSMC: look_bl/look_br flat ragged concat over r of [C(N-r,2), M*K]; cat/res [N-1,K] (all float64)."""
        for t, nm in ((look_bl, "look_bl"), (look_br, "look_br"), (cat, "cat"), (res, "res")):
            _chk(t, F64, nm)
        need = self.K * self.M * sum((self.N - r) * (self.N - r - 1) // 2 for r in range(self.N - 1))
        if look_bl.numel() != need or look_br.numel() != need or cat.numel() != (self.N - 1) * self.K:
            raise ValueError("uniform arrays have the wrong size")
        self._uniforms = (look_bl, look_br, cat, res)
        check(self._lib.vcsmc_sweep_set_uniforms_nested(self._h, _ptr(look_bl), _ptr(look_br), _ptr(cat), _ptr(res)))

    def set_option(self, name: str, value: float) -> None:
        check(self._lib.vcsmc_sweep_set_option(self._h, name.encode(), float(value)))

    def set_allreduce(self, fn) -> None:
        """fn(tensor_f64_1d) sums in place across ranks (site sharding); None removes the hook."""
        if fn is None:
            self._hook = None
            check(self._lib.vcsmc_sweep_set_allreduce(self._h, _lib.ALLREDUCE_FN(0), None))
            return
        ws = self.workspace          # (the closure must not hold `self`: a cycle would delay freeing the workspace)
        base = ws.data_ptr()

        def _cb(_user, buf, count, _stream_):
            try:
                off = buf - base
                fn(ws[off:off + 8 * count].view(F64))
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return -1

        self._hook = _lib.ALLREDUCE_FN(_cb)
        check(self._lib.vcsmc_sweep_set_allreduce(self._h, self._hook, None))

    def set_comm(self, comm) -> None:
        """Particle sharding (``vcsmc_sweep_set_comm``): this rank owns K/world logical particles; ``comm`` (a
        ``phylo_b200.comm.Comm``) provides the per-event all-gather and barrier, the workspaces of all ranks are
        mapped into each other through CUDA IPC."""
        from .comm import PeerMap
        if self.K % comm.world != 0:
            raise ValueError("n_particles=%d is not divisible by the %d ranks" % (self.K, comm.world))
        ws = self.workspace          # (not `self`: see set_allreduce)
        base = ws.data_ptr()
        dev = self.device

        def _cb(_user, op, buf, nbytes, _stream_):
            try:
                if op == _lib.COMM_BARRIER:
                    comm.barrier(dev)
                elif op == _lib.COMM_ALLGATHER:
                    off = buf - base
                    comm.all_gather_inplace(ws[off:off + nbytes * comm.world], nbytes)
                elif op == _lib.COMM_ALLREDUCE:
                    off = buf - base
                    comm.all_reduce(ws[off:off + nbytes].view(F64))
                else:
                    return -1
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return -1

        self._peers = PeerMap(comm, self.workspace)
        self._comm_hook = _lib.COMM_FN(_cb)
        self._comm = comm
        check(self._lib.vcsmc_sweep_set_comm(self._h, comm.rank, comm.world, self._comm_hook, None, self._peers.as_array()))
        comm.barrier(dev)   # every rank has zeroed its flag array before anybody signals
        torch.cuda.synchronize(dev)
        self.retained = False

    # -- execution
    def forward(self, codes: torch.Tensor, lam_l: torch.Tensor, lam_r: torch.Tensor, Q: Optional[torch.Tensor],
                pi: torch.Tensor) -> torch.Tensor:
        _chk(codes, U8, "codes"); _chk(lam_l, F64, "lam_l"); _chk(lam_r, F64, "lam_r"); _chk(pi, F64, "pi")
        if tuple(codes.shape) != (self.N, self.S):
            raise ValueError("codes must be [%d,%d], got %s" % (self.N, self.S, tuple(codes.shape)))
        if not self.jc:
            _chk(Q, F64, "Q")
        self._keepalive = (codes, lam_l, lam_r, Q, pi)
        check(self._lib.vcsmc_sweep_forward(self._h, _ptr(codes), _ptr(lam_l), _ptr(lam_r), _ptr(Q), _ptr(pi), _stream()))
        return self.output("elbo")

    def backward(self, grad_elbo: float = 1.0):
        """Returns (dlam_l[N-1], dlam_r[N-1], dQ[4,4] or None, dpi[4]) of grad_elbo * ELBO."""
        dev = self.device
        dl = torch.empty(self.N - 1, dtype=F64, device=dev)
        dr = torch.empty(self.N - 1, dtype=F64, device=dev)
        dQ = None if self.jc else torch.empty((4, 4), dtype=F64, device=dev)
        dpi = torch.empty(4, dtype=F64, device=dev)
        check(self._lib.vcsmc_sweep_backward(self._h, float(grad_elbo), _ptr(dl), _ptr(dr), _ptr(dQ), _ptr(dpi), _stream()))
        return dl, dr, dQ, dpi

    def profile(self) -> Dict[str, Tuple[float, int]]:
        """{kernel: (total ms, launches)} since set_option('profile', 1); resets.  merge_fwd = the forward's scoring /
        merge kernels, event_kernel = the lazy forward's cooperative bookkeeping kernel (one launch per rank event)."""
        buf = (C.c_double * 8)()
        check(self._lib.vcsmc_sweep_profile(self._h, buf))
        return {"merge_fwd": (buf[0], int(buf[1])), "merge_fwd_recompute": (buf[2], int(buf[3])),
                "merge_bwd": (buf[4], int(buf[5])), ("lookahead" if self.M else "event_kernel"): (buf[6], int(buf[7]))}

    def rem_positions(self):
        """Per rank event r, the uint8 [K, N-r-2] table of kept forest positions (host numpy), see the header."""
        ptr = self._lib.vcsmc_sweep_output(self._h, b"rem_positions")
        base = ptr - self.workspace.data_ptr()
        out, off = [], 0
        for r in range(self.N - 1):
            cnt = self.K * (self.N - r - 2)
            out.append(self.workspace[base + off: base + off + cnt].cpu().numpy().reshape(self.K, self.N - r - 2))
            off += (cnt + 15) // 16 * 16
        return out

    def event_timing(self):
        """[N,16] uint64 %globaltimer stamps (ns) of the lazy forward's event kernel, launch r in row r: entry i is taken
        by CTA 0 at the start of phase i (needs set_option("event_timing", 1) before the forward)."""
        ptr = self._lib.vcsmc_sweep_output(self._h, b"event_timing")
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + 8 * 16 * self.N].view(torch.int64).view(self.N, 16).cpu().numpy()

    def rem_row(self, r: int, k: int):
        """Kept forest positions of ONE particle slot at rank event r (uint8 [N-r-2], host numpy)."""
        ptr = self._lib.vcsmc_sweep_output(self._h, b"rem_positions")
        off = ptr - self.workspace.data_ptr()
        for q in range(r):
            off += (self.K * (self.N - q - 2) + 15) // 16 * 16
        m = self.N - r - 2
        return self.workspace[off + k * m: off + (k + 1) * m].cpu().numpy()

    def check_status(self) -> Dict[str, int]:
        """Synchronises; raises if the device-side status word reports an error (e.g. node pool exhausted)."""
        st = self.output("status").cpu().tolist()
        if st[0] != 0:
            raise _lib.VcsmcError(st[0], "device-side failure during the sweep (status=%s); pool exhausted means the "
                                  "workspace is too small for the number of live nodes" % st)
        return {"peak_pool_slots": st[1], "backward_chunks": st[2], "backward_events_visited": st[3]}

    def output(self, name: str) -> torch.Tensor:
        """A view (no copy) of a result table inside the workspace; valid until the next forward()."""
        ptr = self._lib.vcsmc_sweep_output(self._h, name.encode())
        if not ptr:
            raise KeyError(name)
        N1K = (self.N - 1, self.K)
        shapes = {None: N1K, "K": (self.K,), "N1": (self.N - 1,)}
        if name in _F64_OUT:
            spec, dtype, isz = _F64_OUT[name], F64, 8
        elif name in _I32_OUT:
            spec, dtype, isz = _I32_OUT[name], I32, 4
        else:
            raise KeyError(name)
        shape = spec if isinstance(spec, tuple) else shapes[spec]
        n = 1
        for d in shape:
            n *= d
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + n * isz].view(dtype).view(shape)


class _SweepElbo(torch.autograd.Function):
    """ELBO of one sweep with the hand-written reverse sweep as its gradient."""

    @staticmethod
    def forward(ctx, sweep: Sweep, codes, lam_l, lam_r, Q, pi):
        elbo = sweep.forward(codes, lam_l.detach().contiguous(), lam_r.detach().contiguous(),
                             None if Q is None else Q.detach().contiguous(), pi.detach().contiguous())
        ctx.sweep = sweep
        ctx.has_Q = Q is not None
        return elbo.clone().reshape(())

    @staticmethod
    def backward(ctx, grad):
        # grad stays on the device: the library scales by a host double, so fold it in afterwards
        dl, dr, dQ, dpi = ctx.sweep.backward(1.0)
        g = grad.to(F64)
        return (None, None, dl * g, dr * g, (dQ * g) if (ctx.has_Q and dQ is not None) else None, dpi * g)


def sweep_elbo(sweep: Sweep, codes, lam_l, lam_r, Q, pi) -> torch.Tensor:
    """Differentiable ELBO (vcsmc.py:445) w.r.t. lam_l, lam_r [N-1], Q [4,4], pi [4]."""
    return _SweepElbo.apply(sweep, codes, lam_l, lam_r, Q, pi)


# ------------------------------------------------------------------------------------------
# torch.library registration (torch.ops.vcsmc.*)
# ------------------------------------------------------------------------------------------
def _register():
    lib = torch.library

    @lib.custom_op("vcsmc::transition", mutates_args=())
    def transition(Q: torch.Tensor, t: torch.Tensor, jc: bool) -> torch.Tensor:
        return transition_fwd(None if jc else Q.contiguous(), t.contiguous(), jc)

    @transition.register_fake
    def _(Q, t, jc):
        return t.new_empty((t.numel(), 4, 4))

    def _transition_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1])
        ctx.jc = inputs[2]

    def _transition_bwd(ctx, gP):
        Q, t = ctx.saved_tensors
        if ctx.jc:
            eye = torch.eye(4, dtype=F64, device=gP.device)
            comp = torch.zeros((t.numel(), 16), dtype=F64, device=gP.device)
            comp[:, 0] = (gP * eye).sum(dim=(1, 2))
            comp[:, 1] = (gP * (1 - eye)).sum(dim=(1, 2))
            dt, _ = transition_bwd(None, t.contiguous(), comp, True)
            return None, dt, None
        dt, dQ = transition_bwd(Q.contiguous(), t.contiguous(), gP.contiguous(), False)
        return dQ.sum(dim=0), dt, None

    transition.register_autograd(_transition_bwd, setup_context=_transition_setup)

    @lib.custom_op("vcsmc::propose_pairs", mutates_args=())
    def _propose(u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return propose_pairs(u.contiguous())

    @_propose.register_fake
    def _(u):
        return u.new_empty((u.shape[0], 2), dtype=I32), u.new_empty((u.shape[0], max(u.shape[1] - 2, 0)), dtype=I32)

    @lib.custom_op("vcsmc::resample", mutates_args=())
    def _resample(lw: torch.Tensor, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        idx, lse, ess = resample(lw.contiguous(), u.contiguous())
        return idx, lse.clone(), ess.clone()

    @_resample.register_fake
    def _(lw, u):
        return lw.new_empty(lw.shape, dtype=I32), lw.new_empty(()), lw.new_empty(())

    @lib.custom_op("vcsmc::merge", mutates_args=())
    def merge(L_l: torch.Tensor, L_r: torch.Tensor, P_l: torch.Tensor, P_r: torch.Tensor, pi: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Dense-children merge: L_l, L_r [K,S,4]; P_l, P_r [K,4,4]; returns (new [K,S,4], ell [K])."""
        K, S = L_l.shape[0], L_l.shape[1]
        pool = torch.empty((3 * K, S, 4), dtype=F64, device=L_l.device)
        pool[:K] = L_l
        pool[K:2 * K] = L_r
        ar = torch.arange(K, dtype=I32, device=L_l.device)
        P = torch.cat([P_l.reshape(K, 16), P_r.reshape(K, 16)], dim=1).contiguous()
        ell = merge_fwd(None, pool, ar, ar + K, ar + 2 * K, P, pi.contiguous(), S, False)
        return pool[2 * K:].clone(), ell

    @merge.register_fake
    def _(L_l, L_r, P_l, P_r, pi):
        return L_l.new_empty(L_l.shape), L_l.new_empty((L_l.shape[0],))

    def _merge_setup(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    def _merge_bwd(ctx, g_new, g_ell):
        L_l, L_r, P_l, P_r, pi = ctx.saved_tensors
        K, S = L_l.shape[0], L_l.shape[1]
        dev = L_l.device
        pool = torch.empty((3 * K, S, 4), dtype=F64, device=dev)
        pool[:K] = L_l
        pool[K:2 * K] = L_r
        gpool = torch.zeros((3 * K, S, 4), dtype=F64, device=dev)
        if g_new is not None:
            gpool[2 * K:] = g_new
        ar = torch.arange(K, dtype=I32, device=dev)
        P = torch.cat([P_l.reshape(K, 16), P_r.reshape(K, 16)], dim=1).contiguous()
        coef = (g_ell if g_ell is not None else torch.zeros(K, dtype=F64, device=dev)).contiguous()
        dP = torch.zeros((K, 32), dtype=F64, device=dev)
        dpi = torch.zeros(4, dtype=F64, device=dev)
        merge_bwd(None, pool, gpool, ar, ar + K, ar + 2 * K, P, pi.contiguous(), coef, S, False, dP, dpi)
        return gpool[:K], gpool[K:2 * K], dP[:, :16].reshape(K, 4, 4), dP[:, 16:].reshape(K, 4, 4), dpi

    merge.register_autograd(_merge_bwd, setup_context=_merge_setup)


_register()

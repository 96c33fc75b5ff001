"""Restarted Lanczos ground-state solver on the fused sm_100a vector kernels.

Drop-in for ``quantum_simulation.eigensolver.lanczos_ground_state``
(reference ``eigensolver/lanczos_solver.py:310-358`` -> ``_LanczosEngine.forward`` ``:157-303``): same
signature, same restart / Krylov / PRO / early-stop semantics, same return convention (the vector aliases a
workspace buffer, ``:294``; the energy is the Ritz value, ``:303``).

What changed underneath:
  * the Krylov basis is stored ``k x n`` (rows contiguous) and ``v_j`` *is* row ``j`` -- no ``K[:, j]`` strided
    copies (``:230``), no ``vec_prev <- vec_cur`` copies (``:268-269``);
  * ``vdot`` (``:241``), the two ``add_`` (``:245-247``), ``norm`` (``:259``), the PRO GEMVs (``:254-256``), the
    ``div_`` (``:270``) and the Ritz ``addmv`` (``:293``) are the fused kernels ``qsb_lz_*`` (8 n-element memory
    passes per iteration instead of ~17), with alpha/beta/overlaps kept in a device scalar array;
  * the per-iteration ``beta.item()`` host sync (``:263``) is gone: the whole restart is enqueued and alpha/beta
    are read back once; a breakdown (beta_j < beta_tol) is honoured after the fact by truncating the tridiagonal
    matrix at the same ``j`` the reference would have stopped at, so results are identical.  Only in the converged
    regime -- after a restart that DID end in a breakdown -- the next restart reads beta back every iteration, so the
    matvecs the reference skips are skipped here too (a converged sweep applies H_eff twice per bond, not 8 times);
  * the tiny ``k x k`` ``eigh`` (``:285``) runs on the host in float64.
Control flow stays in Python.  CUDA tensors only -- there is no CPU fallback.

Multi-GPU: ``parallel.lanczos_ground_state_sharded`` runs the same loop on a row shard of
every vector; the partial dot products / squared norms the kernels leave in the device scalar array are summed
with one small all-reduce each (``group`` below).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Tuple

import torch

from . import _lib

__all__ = ["lanczos_ground_state"]

_MAX_KRYLOV = 64
_REAL, _SQRT, _NOSQRT = 1, 2, 4          # flag bits of qsb_lz_dot / qsb_lz_axpy_norm / qsb_lz_combine


class CudaVectorOps:
    """The product vector backend: torch.ops.qsb200.lz_* (csrc/lanczos.cu)."""

    device_type = "cuda"

    def make_scratch(self, device):
        return torch.zeros(_lib.lz_scratch_bytes(), device=device, dtype=torch.uint8)

    def dot(self, X, m, ldx, w, n, scal, out_idx, flags, scratch):
        _lib.ops.lz_dot(X, m, ldx, w, n, scal, out_idx, flags, scratch)

    def axpy_norm(self, X, m, ldx, w, n, scal, coef_idx, norm_idx, flags, scratch):
        _lib.ops.lz_axpy_norm(X, m, ldx, w, n, scal, coef_idx, norm_idx, flags, scratch)

    def scale(self, src, dst, n, scal, idx):
        _lib.ops.lz_scale(src, dst, n, scal, idx)

    def combine(self, X, m, ldx, y, dst, n, scal, norm_idx, flags, scratch):
        _lib.ops.lz_combine(X, m, ldx, y, dst, n, scal, norm_idx, flags, scratch)


_CUDA_OPS = CudaVectorOps()


class _Workspace:
    """Preallocated buffers for one (n, k, device, dtype) -- reference _LanczosWorkspace (:22-69)."""

    def __init__(self, n: int, k: int, device: torch.device, dtype: torch.dtype, vec):
        self.n, self.k = n, k
        self.K = torch.empty(k, max(n, 1), device=device, dtype=dtype)       # Krylov basis, rows contiguous
        self.w = torch.empty(n, device=device, dtype=dtype)
        self.tmp = torch.empty(n, device=device, dtype=dtype)
        self.best = torch.empty(n, device=device, dtype=dtype)               # tmp_n2 of the reference
        # device scalars, (re, im) pairs: [0] beta_{-1}=0, [2j+1] alpha_j, [2j+2] beta_j, then scratch slots
        self.NRM = 2 * k + 1
        self.OV = 2 * k + 2
        self.scal = torch.zeros(3 * k + 4, 2, device=device, dtype=torch.float64)
        self.scratch = vec.make_scratch(device)
        # True after a restart that ended in a breakdown (beta_j < beta_tol): the next restart then reads beta back
        # every iteration, like the reference (:263), and stops at the same j instead of enqueueing all k matvecs
        self.watch_beta = False


_workspaces: Dict[tuple, _Workspace] = {}


def _ensure_ws(n: int, k: int, device: torch.device, dtype: torch.dtype, vec) -> _Workspace:
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
    key = (n, k, device.type, device.index, dtype, id(vec), stream)      # one workspace per (device, stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _Workspace(n, k, device, dtype, vec)
        _workspaces[key] = ws
    return ws


def clear_workspaces() -> None:
    """Free every cached Lanczos workspace (the reference never evicts; at chi >= 2048 one is > 1 GB) and the cached
    matrices of the two-site split's range compression."""
    _workspaces.clear()
    from . import svd
    svd.clear_caches()


def _lanczos_impl(initial_vector, linear_operator, operator_args, num_restarts, krylov_dim, pro, pro_max_reorth,
                  beta_tol, ritz_residual_tol, *, group: Any = None, vec=_CUDA_OPS):
    """The solver loop.  ``group``: torch.distributed process group when every vector is a row shard (None = one
    GPU).  ``vec``: the vector-op backend (the CUDA kernels; tests of the multi-rank host logic inject their own)."""
    sharded = group is not None
    if sharded:
        import torch.distributed as dist

    def reduce_sum(slot: torch.Tensor) -> None:
        if sharded:
            dist.all_reduce(slot, group=group)

    with torch.inference_mode():
        x0 = initial_vector.reshape(-1)
        if not x0.is_contiguous():
            x0 = x0.contiguous()
        x0 = _lib._resolved(x0)
        n = x0.numel()
        k = max(1, int(krylov_dim))
        if k > _MAX_KRYLOV:
            raise ValueError(f"krylov_dim {k} exceeds the supported maximum {_MAX_KRYLOV}")
        ws = _ensure_ws(n, k, x0.device, x0.dtype, vec)
        K, w, scal, scratch = ws.K, ws.w, ws.scal, ws.scratch
        NRM, OV = ws.NRM, ws.OV

        def norm_to(src, slot):
            """scal[slot] = global ||src||."""
            if sharded:
                vec.dot(src, 1, 0, src, n, scal, slot, _REAL, scratch)
                reduce_sum(scal[slot])
                scal[slot].sqrt_()
            else:
                vec.dot(src, 1, 0, src, n, scal, slot, _REAL | _SQRT, scratch)

        # best_vec <- normalize(initial_vector), zero vector re-seeded from the global RNG (:137-155, :209)
        norm_to(x0, NRM)
        if float(scal[NRM, 0].item()) == 0.0:
            ws.tmp.copy_(torch.empty_like(ws.tmp).uniform_(-1, 1))
            x0 = ws.tmp
            norm_to(x0, NRM)
        vec.scale(x0, ws.best, n, scal, NRM)
        best_e = float("inf")
        nflag = _NOSQRT if sharded else 0

        def finish_norm(slot):
            if sharded:
                reduce_sum(scal[slot])
                scal[slot].sqrt_()

        for _restart in range(int(num_restarts)):
            scal.zero_()                                       # alphas, betas <- 0 (:216-222)
            norm_to(ws.best, NRM)                               # vec_cur <- normalize(best_vec) (:214)
            vec.scale(ws.best, K[0], n, scal, NRM)
            watched_stop = None
            for j in range(k):
                v = K[j]
                try:                                            # (:234-238)
                    linear_operator(v, *operator_args, out=w)
                except TypeError:
                    w.copy_(linear_operator(v, *operator_args))
                vec.dot(v, 1, 0, w, n, scal, 2 * j + 1, _REAL, scratch)             # alpha_j = Re<v_j, w> (:241)
                reduce_sum(scal[2 * j + 1])
                do_pro = bool(pro) and j > 0 and int(pro_max_reorth) > 0
                nidx = -1 if do_pro else 2 * j + 2
                if j == 0:                                      # w -= alpha v_j (- beta v_{j-1}), beta_j = |w| (:245-259)
                    vec.axpy_norm(K[0], 1, 0, w, n, scal, 1, nidx, nflag, scratch)
                else:
                    vec.axpy_norm(K[j - 1], 2, n, w, n, scal, 2 * j, nidx, nflag, scratch)
                if do_pro:                                      # block Gram-Schmidt against K[:j] (:250-256)
                    reps = int(pro_max_reorth)
                    for r in range(reps):
                        vec.dot(K[0], j, n, w, n, scal, OV, 0, scratch)
                        reduce_sum(scal[OV:OV + j])
                        vec.axpy_norm(K[0], j, n, w, n, scal, OV, 2 * j + 2 if r == reps - 1 else -1, nflag, scratch)
                finish_norm(2 * j + 2)
                if ws.watch_beta and j < k - 1:                 # converged regime: stop where the reference stops (:263-265)
                    if not (float(scal[2 * j + 2, 0].item()) >= beta_tol):
                        watched_stop = j + 1
                        break
                if j < k - 1:                                   # v_{j+1} = w / beta_j (:268-270)
                    vec.scale(w, K[j + 1], n, scal, 2 * j + 2)

            ab = scal[: 2 * k + 1, 0].cpu()                     # the one host read-back of this restart
            alphas = ab[1::2]
            betas = ab[2::2]                                    # beta_0 .. beta_{k-1}
            c = k
            for j in range(k if watched_stop is None else watched_stop):     # breakdown / last-iteration stop (:263-265)
                if not (float(betas[j]) >= beta_tol) or j == k - 1:
                    c = j + 1
                    break
            if watched_stop is not None:
                c = min(c, watched_stop)
            ws.watch_beta = c < k
            beta_tail = float(betas[c - 1])
            T = torch.zeros(c, c, dtype=torch.float64)
            T.diagonal().copy_(alphas[:c])
            if c > 1:
                T.diagonal(1).copy_(betas[: c - 1])
                T.diagonal(-1).copy_(betas[: c - 1])
            evals, evecs = torch.linalg.eigh(T)                 # (:277-288)
            i0 = int(torch.argmin(evals))
            theta = float(evals[i0])
            y = evecs[:, i0]
            if theta < best_e:                                  # Ritz vector K[:c]^T y, normalised (:291-295)
                vec.combine(K[0], c, n, [float(t) for t in y], ws.tmp, n, scal, NRM, nflag, scratch)
                finish_norm(NRM)
                vec.scale(ws.tmp, ws.best, n, scal, NRM)
                best_e = theta
            if ritz_residual_tol is not None:                   # (:298-301)
                if abs(beta_tail) * abs(float(y[-1])) < ritz_residual_tol:
                    break
        return ws.best, float(best_e)


def lanczos_ground_state(
        initial_vector: torch.Tensor,
        linear_operator: Callable,
        operator_args: Tuple = (),
        num_restarts: int = 2,
        krylov_dim: int = 4,
        *,
        pro: bool = False,
        pro_max_reorth: int = 1,
        beta_tol: float = 1e-12,
        ritz_residual_tol: Optional[float] = None,
) -> Tuple[torch.Tensor, float]:
    """Lowest eigenpair of a Hermitian operator given as ``linear_operator(v, *operator_args, out=w)``.

    Mirrors the reference signature (lanczos_solver.py:310-321).  ``linear_operator`` may be any callable; if it
    does not accept ``out=`` the result is copied into the workspace, exactly as the reference does (:234-238).
    """
    if initial_vector.device.type != "cuda":
        raise RuntimeError("qsb200 lanczos_ground_state runs on CUDA tensors only (no CPU fallback)")
    with _lib.nvtx_range("qsb.lanczos"):
        return _lanczos_impl(initial_vector, linear_operator, operator_args, num_restarts, krylov_dim, pro,
                             pro_max_reorth, beta_tol, ritz_residual_tol)

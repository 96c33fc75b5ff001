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
    matrix at the same ``j`` the reference would have stopped at, so results are identical;
  * the tiny ``k x k`` ``eigh`` (``:285``) runs on the host in float64.
Control flow stays in Python.  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib

__all__ = ["lanczos_ground_state"]

_MAX_KRYLOV = 64


class _Workspace:
    """Preallocated buffers for one (n, k, device, dtype) -- reference _LanczosWorkspace (:22-69)."""

    def __init__(self, n: int, k: int, device: torch.device, dtype: torch.dtype):
        self.n, self.k = n, k
        self.K = torch.empty(k, max(n, 1), device=device, dtype=dtype)       # Krylov basis, rows contiguous
        self.w = torch.empty(n, device=device, dtype=dtype)
        self.tmp = torch.empty(n, device=device, dtype=dtype)
        self.best = torch.empty(n, device=device, dtype=dtype)               # tmp_n2 of the reference
        # device scalars, (re, im) pairs: [0] beta_{-1}=0, [2j+1] alpha_j, [2j+2] beta_j, then scratch slots
        self.NRM = 2 * k + 1
        self.OV = 2 * k + 2
        self.scal = torch.zeros(3 * k + 4, 2, device=device, dtype=torch.float64)
        self.scratch = torch.zeros(_lib.lz_scratch_bytes(), device=device, dtype=torch.uint8)


_workspaces: Dict[tuple, _Workspace] = {}


def _ensure_ws(n: int, k: int, device: torch.device, dtype: torch.dtype) -> _Workspace:
    key = (n, k, device.type, device.index, dtype)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _Workspace(n, k, device, dtype)
        _workspaces[key] = ws
    return ws


def clear_workspaces() -> None:
    """Free every cached Lanczos workspace (the reference never evicts; at chi >= 2048 one is > 1 GB)."""
    _workspaces.clear()


def _normalize_into(ws: _Workspace, src: torch.Tensor, dst: torch.Tensor) -> None:
    ops = _lib.ops
    ops.lz_dot(src, 1, 0, src, ws.n, ws.scal, ws.NRM, 3, ws.scratch)
    ops.lz_scale(src, dst, ws.n, ws.scal, ws.NRM)


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
    ops = _lib.ops
    with torch.inference_mode():
        x0 = initial_vector.reshape(-1)
        if not x0.is_contiguous():
            x0 = x0.contiguous()
        x0 = _lib._resolved(x0)
        n = x0.numel()
        k = max(1, int(krylov_dim))
        if k > _MAX_KRYLOV:
            raise ValueError(f"krylov_dim {k} exceeds the supported maximum {_MAX_KRYLOV}")
        ws = _ensure_ws(n, k, x0.device, x0.dtype)
        K, w, scal = ws.K, ws.w, ws.scal
        NRM, OV = ws.NRM, ws.OV

        # best_vec <- normalize(initial_vector), zero vector re-seeded from the global RNG (:137-155, :209)
        ops.lz_dot(x0, 1, 0, x0, n, scal, NRM, 3, ws.scratch)
        if float(scal[NRM, 0].item()) == 0.0:
            ws.tmp.copy_(torch.empty_like(ws.tmp).uniform_(-1, 1))
            x0 = ws.tmp
            ops.lz_dot(x0, 1, 0, x0, n, scal, NRM, 3, ws.scratch)
        ops.lz_scale(x0, ws.best, n, scal, NRM)
        best_e = float("inf")

        for _restart in range(int(num_restarts)):
            scal.zero_()                                       # alphas, betas <- 0 (:216-222)
            _normalize_into(ws, ws.best, K[0])                  # vec_cur <- normalize(best_vec) (:214)
            for j in range(k):
                v = K[j]
                try:                                            # (:234-238)
                    linear_operator(v, *operator_args, out=w)
                except TypeError:
                    w.copy_(linear_operator(v, *operator_args))
                ops.lz_dot(v, 1, 0, w, n, scal, 2 * j + 1, 1, ws.scratch)           # alpha_j = Re<v_j, w> (:241)
                do_pro = bool(pro) and j > 0 and int(pro_max_reorth) > 0
                nidx = -1 if do_pro else 2 * j + 2
                if j == 0:                                      # w -= alpha v_j (- beta v_{j-1}), beta_j = |w| (:245-259)
                    ops.lz_axpy_norm(K[0], 1, 0, w, n, scal, 1, nidx, ws.scratch)
                else:
                    ops.lz_axpy_norm(K[j - 1], 2, n, w, n, scal, 2 * j, nidx, ws.scratch)
                if do_pro:                                      # block Gram-Schmidt against K[:j] (:250-256)
                    reps = int(pro_max_reorth)
                    for r in range(reps):
                        ops.lz_dot(K[0], j, n, w, n, scal, OV, 0, ws.scratch)
                        ops.lz_axpy_norm(K[0], j, n, w, n, scal, OV, 2 * j + 2 if r == reps - 1 else -1, ws.scratch)
                if j < k - 1:                                   # v_{j+1} = w / beta_j (:268-270)
                    ops.lz_scale(w, K[j + 1], n, scal, 2 * j + 2)

            ab = scal[: 2 * k + 1, 0].cpu()                     # the one host read-back of this restart
            alphas = ab[1::2]
            betas = ab[2::2]                                    # beta_0 .. beta_{k-1}
            c = k
            for j in range(k):                                  # breakdown / last-iteration stop (:263-265)
                if not (float(betas[j]) >= beta_tol) or j == k - 1:
                    c = j + 1
                    break
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
                ops.lz_combine(K[0], c, n, [float(t) for t in y], ws.tmp, n, scal, NRM, ws.scratch)
                ops.lz_scale(ws.tmp, ws.best, n, scal, NRM)
                best_e = theta
            if ritz_residual_tol is not None:                   # (:298-301)
                if abs(beta_tail) * abs(float(y[-1])) < ritz_residual_tol:
                    break
        return ws.best, float(best_e)

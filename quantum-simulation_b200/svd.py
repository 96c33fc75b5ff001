"""Two-site split  theta -> A . s . B  with the reference's truncation rule ("next" row f.1 of SURVEY.md section 8).

Reference: ``DMRG._sweep`` dmrg.py:684-711 -- ``torch.linalg.svd`` of theta viewed as (chi_l*d, d*chi_r), then
``keep = max(1, min(maxdim, #S, searchsorted(cumsum(S^2/sum S^2), 1-cutoff, right=True)+1))`` (:687-695), Schmidt
values renormalised with +eps (:706-710).

On a GPU the dense SVD is what dominates a sweep once the matvec runs at tensor-core speed (cuSOLVER gesvd: 1.56 s for
4096 x 4096 float64 on B200, vs 0.16 s for the 8 matvecs of the same bond step).  For large blocks this module
therefore computes the same factorisation from the Hermitian eigenproblem of the smaller Gram matrix:

    m <= n:  theta theta^H = U diag(lambda) U^H ,  S = sqrt(lambda) ,  rows of  U_k^H theta = S_k Vh_k
    m >  n:  theta^H theta = V diag(lambda) V^H ,  columns of  theta V_k = U_k S_k

followed by a re-orthonormalisation of the derived factor, so BOTH factors are isometries to machine precision
(what the environment updates rely on).  On CUDA the three chi^3-class products (Gram matrix, projection, the small
Gram matrices of the re-orthonormalisation) run on this library's own FP64 tensor-core GEMM core (bit-reproducible,
no cuBLAS), and the re-orthonormalisation is CholeskyQR2 on the row-normalised factor (two rounds of
Gram -> Cholesky -> triangular solve, all GEMM-shaped) instead of Householder QR (21 ms -> 8 ms at 4096 x 2048), with
Householder QR as the fallback whenever a Cholesky factorisation reports a non-positive pivot (rank-deficient theta).
The dense Hermitian eigenproblem itself stays ``torch.linalg.eigh`` (cuSOLVER syevd, 85 ms at 4096^2): see DESIGN.md
for the alternatives that were measured.  Schmidt directions with S_i / S_max below ~sqrt(eps) = 1e-8 are resolved
only to absolute accuracy ~eps*S_max^2/S_i; they carry weight < 1e-16 of the state, far below the 1e-10 energy
parity bar, and the kept/discarded split only depends on them through that weight.  Small blocks (min dim below
``FAST_MIN``) use ``torch.linalg.svd`` exactly as the reference does.
"""
from __future__ import annotations

from typing import Tuple

import torch

__all__ = ["two_site_split", "FAST_MIN"]

FAST_MIN = 256        # min(m, n) from which the Gram/eigh path is used


def _keep_from_spectrum(S: torch.Tensor, maxdim: int, cutoff: float) -> int:
    """The reference's truncation rule, dmrg.py:687-695."""
    keep_max = min(S.numel(), int(maxdim))
    keep_cut = S.numel()
    if cutoff > 0.0:
        S2 = S ** 2
        csum = torch.cumsum(S2 / torch.sum(S2), dim=0)
        keep_cut = int(torch.searchsorted(csum, torch.tensor(1.0 - cutoff, dtype=csum.dtype, device=csum.device),
                                          right=True)) + 1
    return max(1, min(keep_max, keep_cut))


def _fix_phase(Q: torch.Tensor, Rr: torch.Tensor) -> torch.Tensor:
    """Scale the columns of Q so that the diagonal of R becomes real and positive."""
    d = torch.diagonal(Rr)
    mag = d.abs()
    phase = torch.where(mag > 0, d / mag.clamp_min(torch.finfo(mag.dtype).tiny).to(d.dtype), torch.ones_like(d))
    return Q * phase.unsqueeze(0)


USE_CHOLQR = True     # CholeskyQR2 re-orthonormalisation of the derived factor (False: always Householder QR)


def _mm(a: torch.Tensor, b: torch.Tensor, conj_a: bool = False, conj_b: bool = False) -> torch.Tensor:
    """op(a) @ op(b) for 2-D operands with one unit stride each: the library's own GEMM core on CUDA (float64 /
    complex128), plain torch elsewhere (this module is device-agnostic; the CPU tests use the torch branch)."""
    if a.is_cuda and a.dtype in (torch.float64, torch.complex128) and min(a.shape + b.shape) > 0:
        from . import _lib
        a, b = _lib._resolved(a), _lib._resolved(b)
        if a.stride(0) != 1 and a.stride(1) != 1:
            a = a.contiguous()
        if b.stride(0) != 1 and b.stride(1) != 1:
            b = b.contiguous()
        return _lib.gemm(a, b, conjA=conj_a and a.is_complex(), conjB=conj_b and b.is_complex())
    return (a.conj() if conj_a else a) @ (b.conj() if conj_b else b)


def _orthonormal_rows(B: torch.Tensor):
    """Rows of B (k x n, k <= n, mutually orthogonal up to rounding, norms spanning many decades) -> orthonormal rows
    spanning the same flag of subspaces (row i stays a combination of rows <= i, positive on itself: the gauge of a
    QR factorisation of B^H with a real positive diagonal).  CholeskyQR2 on the row-normalised matrix; None if a
    Cholesky factorisation breaks down (the caller then uses Householder QR)."""
    nrm = torch.linalg.norm(B, dim=1)
    if not bool((nrm > 0).all()):
        return None
    X = B / nrm.unsqueeze(1).to(B.dtype)
    for _ in range(2):
        Gm = _mm(X, X.T, conj_b=True)                        # k x k, = I + small
        Gm = 0.5 * (Gm + Gm.conj().T)
        Lc, info = torch.linalg.cholesky_ex(Gm)
        if int(info) != 0:
            return None
        X = torch.linalg.solve_triangular(Lc, X, upper=False)     # X <- Lc^{-1} X : rows orthonormal
    return X


def _gram_split(theta: torch.Tensor, maxdim: int, cutoff: float):
    m, n = theta.shape
    left = m <= n
    G = _mm(theta, theta.T, conj_b=True) if left else _mm(theta.T, theta, conj_a=True)
    G = 0.5 * (G + G.conj().T)                             # exactly Hermitian input for the eigensolver
    lam, vecs = torch.linalg.eigh(G)                       # ascending
    lam = torch.flip(lam, dims=(0,)).clamp_min(0)
    S = torch.sqrt(lam)
    keep = _keep_from_spectrum(S, maxdim, cutoff)
    top = torch.flip(vecs[:, vecs.shape[1] - keep:], dims=(1,))        # leading `keep` eigenvectors, descending
    # the derived factor: rows of  U_k^H theta = S_k Vh_k  (left) /  (theta V_k)^H = S_k U_k^H  (right)
    B = _mm(top.T, theta, conj_a=True) if left else _mm(top.T, theta.T, conj_a=True, conj_b=True)
    Q = _orthonormal_rows(B) if USE_CHOLQR else None
    if Q is None:
        Qc, Rr = torch.linalg.qr(B.conj().T)                # Householder fallback, same gauge
        Q = _fix_phase(Qc, Rr).conj().T
    if left:
        U, Vh = top, Q                                      # (m, keep) orthonormal columns, (keep, n) orthonormal rows
    else:
        U, Vh = Q.conj().T, top.conj().T
    return U.resolve_conj().contiguous(), S, Vh.resolve_conj().contiguous(), keep


def two_site_split(theta: torch.Tensor, maxdim: int, cutoff: float, *, fast_min: int = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """Returns ``(U[:, :keep], S_all, Vh[:keep, :], keep)`` for the 2-D block ``theta`` (chi_l*d, d*chi_r).
    ``S_all`` is the full descending spectrum (the caller records the discarded weight from ``S_all[keep:]``)."""
    fm = FAST_MIN if fast_min is None else fast_min
    if min(theta.shape) >= fm:
        if theta.is_cuda:
            from ._lib import nvtx_range
            with nvtx_range("qsb.two_site_split"):
                return _gram_split(theta, maxdim, cutoff)
        return _gram_split(theta, maxdim, cutoff)
    U, S, Vh = torch.linalg.svd(theta, full_matrices=False)         # dmrg.py:685
    keep = _keep_from_spectrum(S, maxdim, cutoff)
    return U[:, :keep].resolve_conj(), S, Vh[:keep, :].resolve_conj(), keep

"""Two-site split  theta -> A . s . B  with the reference's truncation rule ("next" row f.1 of SURVEY.md section 8).

Reference: ``DMRG._sweep`` dmrg.py:684-711 -- ``torch.linalg.svd`` of theta viewed as (chi_l*d, d*chi_r), then
``keep = max(1, min(maxdim, #S, searchsorted(cumsum(S^2/sum S^2), 1-cutoff, right=True)+1))`` (:687-695), Schmidt
values renormalised with +eps (:706-710).

On a GPU the dense SVD is what dominates a sweep once the matvec runs at tensor-core speed (cuSOLVER gesvd: 1.56 s for
4096 x 4096 float64 on B200, vs 0.16 s for the 8 matvecs of the same bond step).  For large blocks this module
therefore computes the same factorisation from the Hermitian eigenproblem of the smaller Gram matrix:

    m <= n:  theta theta^H = U diag(lambda) U^H ,  S = sqrt(lambda) ,  rows of  U_k^H theta = S_k Vh_k
    m >  n:  theta^H theta = V diag(lambda) V^H ,  columns of  theta V_k = U_k S_k

followed by a QR re-orthonormalisation of the derived factor, so BOTH factors are isometries to machine precision
(what the environment updates rely on).  Schmidt directions with S_i / S_max below ~sqrt(eps) = 1e-8 are resolved
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


def _gram_split(theta: torch.Tensor, maxdim: int, cutoff: float):
    m, n = theta.shape
    left = m <= n
    G = theta @ theta.conj().T if left else theta.conj().T @ theta
    lam, vecs = torch.linalg.eigh(G)                       # ascending
    lam = torch.flip(lam, dims=(0,)).clamp_min(0)
    S = torch.sqrt(lam)
    keep = _keep_from_spectrum(S, maxdim, cutoff)
    top = torch.flip(vecs[:, vecs.shape[1] - keep:], dims=(1,))        # leading `keep` eigenvectors, descending
    if left:
        U = top                                             # (m, keep), orthonormal columns
        B = U.conj().T @ theta                              # (keep, n) = S_k Vh_k
        Q, Rr = torch.linalg.qr(B.conj().T)                 # (n, keep)
        Vh = _fix_phase(Q, Rr).conj().T
    else:
        Vh = top.conj().T                                   # (keep, n), orthonormal rows
        Cm = theta @ top                                    # (m, keep) = U_k S_k
        Q, Rr = torch.linalg.qr(Cm)
        U = _fix_phase(Q, Rr)
    return U.resolve_conj(), S, Vh.resolve_conj(), keep      # materialise lazy-conj views once


def two_site_split(theta: torch.Tensor, maxdim: int, cutoff: float, *, fast_min: int = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """Returns ``(U[:, :keep], S_all, Vh[:keep, :], keep)`` for the 2-D block ``theta`` (chi_l*d, d*chi_r).
    ``S_all`` is the full descending spectrum (the caller records the discarded weight from ``S_all[keep:]``)."""
    fm = FAST_MIN if fast_min is None else fast_min
    if min(theta.shape) >= fm:
        return _gram_split(theta, maxdim, cutoff)
    U, S, Vh = torch.linalg.svd(theta, full_matrices=False)         # dmrg.py:685
    keep = _keep_from_spectrum(S, maxdim, cutoff)
    return U[:, :keep].resolve_conj(), S, Vh[:keep, :].resolve_conj(), keep

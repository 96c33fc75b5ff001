"""Matrix product state container: host-side mirror of ``quantum_simulation/mps.py``.

Input provider for the DMRG hot path: same constructor, ``A`` / ``B`` / ``sWeight``
ParameterLists, ``ortho_center``, ``position`` and ``random_mps`` as the reference (mps.py:21-168).  It is plain
torch (one-time O(N d chi^3) QR sweep); the kernels of this package only consume its tensors.

Parity note (hard part vi): CPU and CUDA generators differ, so parity runs build the MPS with
``device='cpu', seed=...`` -- bit-identical to the reference's -- and then move it with ``.to('cuda')``.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
import torch.nn as nn

__all__ = ["MPS", "random_mps"]


def random_mps(n_sites: int, d: int, chi: int, *, dtype, device, seed: Optional[int] = None) -> List[torch.Tensor]:
    """Site tensors with uniform[0,1) entries; bond k is min(chi, d^k, d^(N-k))  (mps.py:154-168)."""
    gen = torch.Generator(device=device).manual_seed(seed) if seed is not None else None
    sites: List[torch.Tensor] = []
    left = 1
    for k in range(n_sites):
        cap_right = d ** (n_sites - k - 1)                  # what the sites to the right can support
        right = min(chi, left * d, cap_right) if k > 0 else min(chi, d)
        sites.append(torch.rand(left, d, right, generator=gen, device=device, dtype=dtype))
        left = right
    return sites


def _frozen(t: torch.Tensor) -> nn.Parameter:
    return nn.Parameter(t, requires_grad=False)


def _absorb_right_factor(site: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """site (l, d, r) = R^T-absorbing split: returns (right-isometric site, the l x l' factor to push left).
    QR of the transposed (l, d*r) matrix, as the reference does (mps.py:52-56, 124-128)."""
    l, d, r = site.shape
    Q, Rf = torch.linalg.qr(site.reshape(l, d * r).T)
    return Q.T.reshape(Q.shape[1], d, r), Rf.T


class MPS(nn.Module):
    """Random right-canonical MPS with the orthogonality centre at site 0 (mps.py:21-70)."""

    def __init__(self, Nsites: int, chid: int, initial_bond_dim: int, device: Any, dtype: torch.dtype,
                 seed: Optional[int] = None, *, rescale_carry: bool = False):
        """Same positional signature as the reference.  ``rescale_carry=True`` (an extension, off by default so that the
        tensors stay identical to the reference's) normalises the carry of the QR sweep at every site: the state is the
        same ray, but long chains at large chi no longer end with a zero centre tensor (see the warning below)."""
        super().__init__()
        self.Nsites, self.chid = Nsites, chid
        self.device, self.dtype = device, dtype
        self.ortho_center = 0
        print("Initializing MPS as a right-canonical random state...")
        sites = random_mps(Nsites, chid, initial_bond_dim, dtype=dtype, device=device, seed=seed)
        for p in reversed(range(1, Nsites)):                # right-to-left QR sweep
            sites[p], carry = _absorb_right_factor(sites[p])
            if rescale_carry:
                carry = carry / torch.linalg.norm(carry)
            sites[p - 1] = torch.einsum("lsc,ck->lsk", sites[p - 1], carry)
        sites[0] = sites[0] / torch.linalg.norm(sites[0])
        # Same arithmetic as the reference (mps.py:57-60) -- including its failure mode: the carry is never rescaled,
        # uniform[0,1) sites have a rank-1 mean component of norm ~0.5*sqrt(l*d*r), so for long chains at large chi
        # (L=100, chi=2048: ~1e280 at site 0) the SQUARED norm overflows, linalg.norm returns inf and site 0 becomes
        # exactly zero.  DMRG still works -- the first Lanczos call re-seeds its zero start vector
        # (lanczos_solver.py:150-153) -- but from the GLOBAL RNG, so such a run repeats only under torch.manual_seed.
        if not bool(torch.isfinite(sites[0]).all()) or not bool(sites[0].any()):
            import warnings
            warnings.warn("MPS: the un-normalised canonicalisation carry over/underflowed (site 0 is zero or "
                          "non-finite), exactly as in the reference at this size; the first DMRG update will start from "
                          "a vector drawn from the global RNG -- call torch.manual_seed(...) for a repeatable run, or "
                          "build the state with rescale_carry=True",
                          RuntimeWarning, stacklevel=2)
        ones = [torch.ones(1, device=device, dtype=dtype) for _ in range(Nsites + 1)]
        self._A = nn.ParameterList(_frozen(t) for t in sites)
        self._B = nn.ParameterList(_frozen(t.clone()) for t in sites)
        self._sWeight = nn.ParameterList(_frozen(t) for t in ones)
        print(f"Initialization complete. Ortho center at site {self.ortho_center}.")

    A = property(lambda self: self._A, doc="left-canonical site tensors (ParameterList)")
    B = property(lambda self: self._B, doc="right-canonical site tensors (ParameterList)")
    sWeight = property(lambda self: self._sWeight, doc="Schmidt values on the bonds (ParameterList)")

    def _set_site(self, p: int, t: torch.Tensor) -> None:
        self._A[p] = _frozen(t)
        self._B[p] = _frozen(t.clone())

    def position(self, site_idx: int) -> None:
        """Move the orthogonality centre: SVD steps to the right, QR steps to the left (mps.py:87-145)."""
        if site_idx < 0 or site_idx >= self.Nsites:
            raise ValueError(f"site_idx {site_idx} out of bounds for Nsites={self.Nsites}")
        print(f"\nMoving orthogonality center from {self.ortho_center} to {site_idx}...")
        while self.ortho_center < site_idx:
            p = self.ortho_center
            l, d, r = self._A[p].shape
            U, S, Vh = torch.linalg.svd(self._A[p].reshape(l * d, r), full_matrices=False)
            self._set_site(p, U.reshape(l, d, U.shape[1]))
            schmidt = S / torch.linalg.norm(S)
            self._sWeight[p + 1] = _frozen(schmidt)
            self._set_site(p + 1, torch.einsum("k,kl,lds->kds", schmidt.to(Vh.dtype), Vh, self._B[p + 1]))
            self.ortho_center = p + 1
        while self.ortho_center > site_idx:
            p = self.ortho_center
            iso, carry = _absorb_right_factor(self._A[p])
            self._set_site(p, iso)
            self._set_site(p - 1, torch.einsum("lsc,ck->lsk", self._A[p - 1], carry))
            self.ortho_center = p - 1
        c = self.ortho_center
        nrm = torch.linalg.norm(self._A[c])
        self._A[c].data /= nrm
        self._B[c].data /= nrm
        print(f"Orthogonality center is now at site {self.ortho_center}.")

    def to(self, *args, **kwargs):
        moved = super().to(*args, **kwargs)
        if len(self._A) > 0:
            self.device = self._A[0].device
        return moved

    def __repr__(self) -> str:
        bonds = [self._A[0].shape[0]] + [t.shape[2] for t in self._A]
        return f"MPS(Nsites={self.Nsites}, chid={self.chid}, ortho_center={self.ortho_center}, bond_dims={bonds})"

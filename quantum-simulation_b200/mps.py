"""Matrix product state container: host-side mirror of ``quantum_simulation/mps.py``.

Input provider for the DMRG hot path: same constructor, ``A`` / ``B`` / ``sWeight``
ParameterLists, ``ortho_center``, ``position`` and ``random_mps`` as the reference (mps.py:21-168).  It is plain
torch (one-time O(N d chi^3) QR sweep); the kernels of this package only consume its tensors.

Parity note (hard part vi): CPU and CUDA generators differ, so parity runs build the MPS with
``device='cpu', seed=...`` -- bit-identical to the reference's -- and then move it with ``.to('cuda')``.
"""
from __future__ import annotations

from typing import Any, List, Optional

import torch
import torch.nn as nn

__all__ = ["MPS", "random_mps"]


def random_mps(n_sites: int, d: int, chi: int, *, dtype, device, seed: Optional[int] = None) -> List[torch.Tensor]:
    """Uniform[0,1) site tensors with bond dims min(chi, d^k, d^(N-k))  (mps.py:154-168)."""
    gen = None
    if seed is not None:
        gen = torch.Generator(device=device).manual_seed(seed)

    def rand(*shape):
        return torch.rand(*shape, generator=gen, device=device, dtype=dtype)

    tensors = [rand(1, d, min(chi, d))]
    for k in range(1, n_sites):
        left = tensors[k - 1].shape[2]
        right = min(min(chi, left * d), d ** (n_sites - k - 1))
        tensors.append(rand(left, d, right))
    return tensors


def _as_param(t: torch.Tensor) -> nn.Parameter:
    return nn.Parameter(t, requires_grad=False)


class MPS(nn.Module):
    """Random right-canonical MPS with the orthogonality centre at site 0 (mps.py:21-70)."""

    def __init__(self, Nsites: int, chid: int, initial_bond_dim: int, device: Any, dtype: torch.dtype,
                 seed: Optional[int] = None):
        super().__init__()
        self.Nsites, self.chid, self.device, self.dtype = Nsites, chid, device, dtype
        self.ortho_center = 0
        print("Initializing MPS as a right-canonical random state...")
        T = random_mps(Nsites, chid, initial_bond_dim, dtype=dtype, device=device, seed=seed)
        for p in range(Nsites - 1, 0, -1):                      # QR sweep from the right (mps.py:52-58)
            chil, d, chir = T[p].shape
            Q, R = torch.linalg.qr(T[p].reshape(chil, d * chir).T)
            T[p] = Q.T.reshape(Q.shape[1], d, chir)
            T[p - 1] = torch.einsum("lsc,ck->lsk", T[p - 1], R.T)
        T[0] = T[0] / torch.linalg.norm(T[0])
        self._A = nn.ParameterList([_as_param(t) for t in T])
        self._B = nn.ParameterList([_as_param(t.clone()) for t in T])
        self._sWeight = nn.ParameterList([_as_param(torch.ones(1, device=device, dtype=dtype))
                                          for _ in range(Nsites + 1)])
        print(f"Initialization complete. Ortho center at site {self.ortho_center}.")

    @property
    def A(self) -> nn.ParameterList:
        return self._A

    @property
    def B(self) -> nn.ParameterList:
        return self._B

    @property
    def sWeight(self) -> nn.ParameterList:
        return self._sWeight

    def position(self, site_idx: int) -> None:
        """Move the orthogonality centre by SVD (rightwards) / QR (leftwards) steps (mps.py:87-145)."""
        if not (0 <= site_idx < self.Nsites):
            raise ValueError(f"site_idx {site_idx} out of bounds for Nsites={self.Nsites}")
        print(f"\nMoving orthogonality center from {self.ortho_center} to {site_idx}...")
        while self.ortho_center < site_idx:
            p = self.ortho_center
            chil, d, chir = self._A[p].shape
            U, S, Vh = torch.linalg.svd(self._A[p].reshape(chil * d, chir), full_matrices=False)
            self._A[p] = _as_param(U.reshape(chil, d, U.shape[1]))
            self._B[p] = _as_param(self._A[p].clone())
            sv = S / torch.linalg.norm(S)
            self._sWeight[p + 1] = _as_param(sv)
            nxt = torch.einsum("k,kl,lds->kds", sv.to(Vh.dtype), Vh, self._B[p + 1])
            self._A[p + 1] = _as_param(nxt)
            self._B[p + 1] = _as_param(nxt.clone())
            self.ortho_center += 1
        while self.ortho_center > site_idx:
            p = self.ortho_center
            chil, d, chir = self._A[p].shape
            Q, R = torch.linalg.qr(self._A[p].reshape(chil, d * chir).T)
            self._A[p] = _as_param(Q.T.reshape(Q.shape[1], d, chir))
            self._B[p] = _as_param(self._A[p].clone())
            prev = torch.einsum("lsc,ck->lsk", self._A[p - 1], R.T)
            self._A[p - 1] = _as_param(prev)
            self._B[p - 1] = _as_param(prev.clone())
            self.ortho_center -= 1
        norm = torch.linalg.norm(self._A[self.ortho_center])
        self._A[self.ortho_center].data /= norm
        self._B[self.ortho_center].data /= norm
        print(f"Orthogonality center is now at site {self.ortho_center}.")

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        for t in self._A:
            self.device = t.device
            break
        return out

    def __repr__(self) -> str:
        bonds = [self._A[0].shape[0]] + [t.shape[2] for t in self._A]
        return f"MPS(Nsites={self.Nsites}, chid={self.chid}, ortho_center={self.ortho_center}, bond_dims={bonds})"

"""Environment-based MPO expectation values ("next" row f.4 of SURVEY.md section 8).

The reference's ``energy_variance`` (dmrg_observables.py:539-582) contracts the MPS to a dense ``d^N`` vector and the
MPO to a dense matrix -- "a validation baseline for small systems"; its docstring defers large calculations to
"optimized environment contractions".  These are those contractions, on the same left-environment kernel the sweep
uses (``qsb_env_left``): ``<psi|O|psi>`` is the chain  L <- L . conj(T_p) . M_p . T_p  over all sites, so the cost is
O(N W d chi^3) instead of O(d^N), and it is an independent check of a DMRG run's convergence at BASELINE sizes:

  energy_expectation(mps, mpo)    <psi|H|psi> / <psi|psi>
  energy_variance(mps, mpo)       <H^2> - <H>^2 with H^2 as the site-wise squared MPO (bond W^2)

Same call signature as the reference (``mps, mpo, *, form="B"``): the state is the product of the ``B`` (or ``A``)
tensors exactly as ``_dense_state_from_mps`` (dmrg_observables.py:236-263) defines it.  CUDA tensors only.
"""
from __future__ import annotations

from typing import Any, List, Sequence

import torch

from .dmrg import LeftEnvUpdate

__all__ = ["mpo_expectation", "energy_expectation", "energy_variance", "squared_mpo"]


def _site_tensors(mps: Any, form: str) -> List[torch.Tensor]:
    if form not in ("A", "B"):
        raise ValueError("form must be 'A' or 'B'.")
    tensors = [t.detach() for t in getattr(mps, form)]
    if len(tensors) != mps.Nsites:
        raise ValueError("MPS tensor count does not match mps.Nsites.")
    return tensors


def _mpo_list(mpo: Any, n: int, device, dtype) -> List[torch.Tensor]:
    if hasattr(mpo, "get_mpo"):
        mpo = mpo.get_mpo()
    if not isinstance(mpo, (list, tuple)):
        raise TypeError("mpo must be a per-site list of (Wl, Wr, d, d) tensors or a builder with get_mpo().")
    if len(mpo) != n:
        raise ValueError(f"MPO has {len(mpo)} sites, the MPS has {n}.")
    return [torch.as_tensor(m).to(device=device, dtype=dtype).contiguous() for m in mpo]


def squared_mpo(mpo: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Site-wise product MPO of H.H: bond (w1, w2), physical legs contracted through the middle."""
    out = []
    for M in mpo:
        Wl, Wr, d, _ = M.shape
        out.append(torch.einsum("abst,cdtu->acbdsu", M, M).reshape(Wl * Wl, Wr * Wr, d, d).contiguous())
    return out


def mpo_expectation(mps: Any, mpo: Sequence[torch.Tensor], *, form: str = "B") -> torch.Tensor:
    """<psi| O |psi> (not normalised) by left-environment contraction; complex scalar tensor on the device."""
    tensors = _site_tensors(mps, form)
    dev, dt = tensors[0].device, tensors[0].dtype
    if dev.type != "cuda":
        raise RuntimeError("qsb200 observables run on CUDA tensors only (no CPU fallback)")
    ops = _mpo_list(mpo, len(tensors), dev, dt)
    upd = LeftEnvUpdate()
    L = torch.ones(ops[0].shape[0], tensors[0].shape[0], tensors[0].shape[0], device=dev, dtype=dt)
    if tensors[0].shape[0] != 1:
        L = torch.eye(tensors[0].shape[0], device=dev, dtype=dt).expand(ops[0].shape[0], -1, -1).contiguous()
    for T, M in zip(tensors, ops):
        L = upd(L, M, T)
    # open right boundary: trace over the (normally one-dimensional) last MPS bond, sum over the last MPO bond
    return torch.diagonal(L, dim1=1, dim2=2).sum()


def _identity_mpo(n: int, d: int, device, dtype) -> List[torch.Tensor]:
    eye = torch.eye(d, device=device, dtype=dtype).reshape(1, 1, d, d)
    return [eye] * n


def energy_expectation(mps: Any, mpo: Any, *, form: str = "B") -> torch.Tensor:
    """<psi|H|psi> / <psi|psi>, real scalar tensor."""
    tensors = _site_tensors(mps, form)
    dev, dt = tensors[0].device, tensors[0].dtype
    H = _mpo_list(mpo, len(tensors), dev, dt)
    norm = mpo_expectation(mps, _identity_mpo(len(tensors), H[0].shape[2], dev, dt), form=form).real
    if float(norm) == 0.0:
        raise ValueError("Cannot compute an expectation value for a zero-norm state.")
    return mpo_expectation(mps, H, form=form).real / norm


def energy_variance(mps: Any, mpo: Any, *, form: str = "B") -> torch.Tensor:
    """Var(H) = <H^2> - <H>^2 by environment contraction (reference semantics: dmrg_observables.py:539-582,
    including the clamp of tiny negative rounding noise to zero)."""
    tensors = _site_tensors(mps, form)
    dev, dt = tensors[0].device, tensors[0].dtype
    H = _mpo_list(mpo, len(tensors), dev, dt)
    norm = mpo_expectation(mps, _identity_mpo(len(tensors), H[0].shape[2], dev, dt), form=form).real
    if float(norm) == 0.0:
        raise ValueError("Cannot compute energy variance for a zero-norm state.")
    energy = mpo_expectation(mps, H, form=form).real / norm
    h2 = mpo_expectation(mps, squared_mpo(H), form=form).real / norm
    variance = h2 - energy.pow(2)
    eps = torch.finfo(variance.dtype).eps
    tol = 100 * eps * max(1.0, abs(float(h2)), float(energy) ** 2)
    if float(variance) < 0 and abs(float(variance)) < tol:
        variance = torch.zeros((), device=dev, dtype=variance.dtype)
    return variance

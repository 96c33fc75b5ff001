"""Environment-based MPO expectation values ("next" row f.4 of SURVEY.md section 8).

The reference's ``energy_variance`` (dmrg_observables.py:539-582) contracts the MPS to a dense ``d^N`` vector and the
MPO to a dense matrix -- "a validation baseline for small systems"; its docstring defers large calculations to
"optimized environment contractions".  These are those contractions, on the same left-environment kernel the sweep
uses (``qsb_env_left``): ``<psi|O|psi>`` is the chain  L <- L . conj(T_p) . M_p . T_p  over all sites, so the cost is
O(N W d chi^3) instead of O(d^N), and it is an independent check of a DMRG run's convergence at BASELINE sizes:

  energy_expectation(mps, mpo)    <psi|H|psi> / <psi|psi>
  energy_variance(mps, mpo)       <H^2> - <H>^2 with H^2 as the site-wise squared MPO (bond W^2)

and, on the same kernels with bond-1 "MPO" tensors (a local operator is a (1, 1, d, d) MPO site), every
product-operator observable the reference computes from the dense vector (dmrg_observables.py:352-449,
dmrg_advanced_observables.py:184-565):

  local_expectation, two_point_correlation, string_order_parameter, many_body_polarization   one operator string
  structure_factor, correlation_length      the whole two-point matrix / one row of it, O(N^2) / O(N) transfer steps
  ground_state_fidelity                      |<a|b>| of two MPS (mixed bra/ket transfer steps)
  entanglement_entropy, entanglement_spectrum                                              from the Schmidt weights

Same call signature as the reference (``mps, mpo, *, form="B"``): the state is the product of the ``B`` (or ``A``)
tensors exactly as ``_dense_state_from_mps`` (dmrg_observables.py:236-263) defines it.  CUDA tensors only.
"""
from __future__ import annotations

from typing import Any, List, Sequence

import torch

from typing import Dict, Optional, Tuple, Union

from .dmrg import LeftEnvUpdate, RightEnvUpdate

__all__ = ["mpo_expectation", "energy_expectation", "energy_variance", "squared_mpo", "local_expectation",
           "two_point_correlation", "string_order_parameter", "many_body_polarization", "structure_factor",
           "correlation_length", "ground_state_fidelity", "entanglement_entropy", "entanglement_spectrum",
           "product_expectation", "ground_state_energy", "energy_density", "excitation_gap", "truncation_errors"]


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


# ------------------------------------------------------------------------------------------------------------------
# product-operator observables on the transfer-operator chain
# ------------------------------------------------------------------------------------------------------------------
def _operator(op: Any, d: int, device, dtype) -> torch.Tensor:
    """Validation of a local operator as in the reference (dmrg_observables.py:281-304): a (d, d) matrix."""
    t = torch.as_tensor(op)
    if t.dim() != 2 or t.shape[0] != d or t.shape[1] != d:
        raise ValueError(f"operator must have shape ({d}, {d}).")
    return t.to(device=device)


def _common_dtype(*dts: torch.dtype) -> torch.dtype:
    out = dts[0]
    for dt in dts[1:]:
        out = torch.promote_types(out, dt)
    if out not in (torch.float64, torch.complex128):
        out = torch.complex128 if out.is_complex else torch.float64
    return out


class _Chain:
    """<psi| prod_k O_k |psi> for ONE MPS as a chain of environment updates with bond-1 MPO sites.

    ``E`` matrices are (1, chi_bra, chi_ket) environments.  ``left(i)`` / ``right(j)`` are the identity-operator
    environments of sites < i / > j, computed once and cached, so any operator string costs only the sites it spans."""

    def __init__(self, mps: Any, form: str, dtype: Optional[torch.dtype] = None):
        tensors = _site_tensors(mps, form)
        if tensors[0].device.type != "cuda":
            raise RuntimeError("qsb200 observables run on CUDA tensors only (no CPU fallback)")
        dt = _common_dtype(tensors[0].dtype, dtype) if dtype is not None else _common_dtype(tensors[0].dtype)
        self.T = [t.to(dt).contiguous() for t in tensors]
        self.N, self.d, self.dev, self.dt = len(tensors), tensors[0].shape[1], tensors[0].device, dt
        self._l, self._r = LeftEnvUpdate(), RightEnvUpdate()
        self._eye = torch.eye(self.d, device=self.dev, dtype=dt).reshape(1, 1, self.d, self.d)
        self._left: Dict[int, torch.Tensor] = {0: self._boundary(self.T[0].shape[0])}
        self._right: Dict[int, torch.Tensor] = {self.N - 1: self._boundary(self.T[-1].shape[2])}

    def _boundary(self, chi: int) -> torch.Tensor:
        return torch.eye(chi, device=self.dev, dtype=self.dt).reshape(1, chi, chi).contiguous()

    def op(self, O: Any) -> torch.Tensor:
        return _operator(O, self.d, self.dev, self.dt).to(self.dt).reshape(1, 1, self.d, self.d).contiguous()

    def step(self, E: torch.Tensor, site: int, M: Optional[torch.Tensor] = None) -> torch.Tensor:
        """E (environment of the sites < ``site``) -> environment of the sites <= ``site`` with operator M there."""
        return self._l(E, self._eye if M is None else M, self.T[site])

    def step_back(self, F: torch.Tensor, site: int, M: Optional[torch.Tensor] = None) -> torch.Tensor:
        """F (environment of the sites > ``site``) -> environment of the sites >= ``site``."""
        return self._r(self._eye if M is None else M, F, self.T[site])

    def left(self, i: int) -> torch.Tensor:
        k = max(j for j in self._left if j <= i)
        while k < i:
            self._left[k + 1] = self.step(self._left[k], k)
            k += 1
        return self._left[i]

    def right(self, j: int) -> torch.Tensor:
        k = min(i for i in self._right if i >= j)
        while k > j:
            self._right[k - 1] = self.step_back(self._right[k], k)
            k -= 1
        return self._right[j]

    def close(self, E: torch.Tensor, last_site: int) -> torch.Tensor:
        """Scalar: E covers the sites <= last_site, the identity environment the rest."""
        return (E * self.right(last_site)).sum()

    def norm2(self) -> torch.Tensor:
        E = self.left(self.N - 1)
        return self.close(self.step(E, self.N - 1), self.N - 1)

    def string(self, ops: Dict[int, torch.Tensor]) -> torch.Tensor:
        """<psi| prod_{k in ops} O_k |psi> (not normalised); ``ops`` maps site -> (1, 1, d, d) operator site."""
        lo, hi = min(ops), max(ops)
        E = self.left(lo)
        for k in range(lo, hi + 1):
            E = self.step(E, k, ops.get(k))
        return self.close(E, hi)


def _check_site(mps: Any, site: int) -> None:
    if not isinstance(site, int) or isinstance(site, bool) or site < 0 or site >= mps.Nsites:
        raise ValueError(f"site index {site!r} is out of range for an MPS with {mps.Nsites} sites.")


def _finish(value: torch.Tensor, norm2: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    if float(norm2.abs()) == 0.0:
        raise ValueError("Cannot compute expectation value for a zero-norm state.")
    out = value / norm2
    return out if dt.is_complex else out.real


def product_expectation(mps: Any, operators: Dict[int, Any], *, form: str = "B") -> torch.Tensor:
    """< prod_k O_k > / <psi|psi> for local operators on distinct sites (``operators``: site -> (d, d) matrix)."""
    if not operators:
        raise ValueError("product_expectation needs at least one operator.")
    for k in operators:
        _check_site(mps, k)
    d = mps.chid
    dt = _common_dtype(getattr(mps, form)[0].dtype if form in ("A", "B") else torch.float64,
                       *[torch.as_tensor(o).dtype for o in operators.values()])
    ch = _Chain(mps, form, dt)
    ops = {k: ch.op(_operator(o, d, ch.dev, dt)) for k, o in operators.items()}
    return _finish(ch.string(ops), ch.norm2(), dt)


def local_expectation(mps: Any, operator: Any, site: int, *, form: str = "B") -> torch.Tensor:
    """<O_site> (reference: dmrg_observables.py:352-389, dense) by environment contraction."""
    _check_site(mps, site)
    return product_expectation(mps, {site: operator}, form=form)


def two_point_correlation(mps: Any, op_i: Any, i: int, op_j: Any, j: int, *, form: str = "B") -> torch.Tensor:
    """<O_i O_j> (dmrg_observables.py:392-449); for i == j the local product ``op_i @ op_j`` is measured."""
    _check_site(mps, i)
    _check_site(mps, j)
    if i == j:
        a, b = torch.as_tensor(op_i), torch.as_tensor(op_j)
        dt = torch.promote_types(a.dtype, b.dtype)
        return local_expectation(mps, a.to(dt) @ b.to(dt), i, form=form)
    return product_expectation(mps, {i: op_i, j: op_j}, form=form)


def string_order_parameter(mps: Any, left_operator: Any, string_operator: Any, right_operator: Any, i: int, j: int,
                           *, form: str = "B") -> torch.Tensor:
    """< O_left(i) prod_{i<k<j} O_string(k) O_right(j) >  (dmrg_advanced_observables.py:417-468)."""
    _check_site(mps, i)
    _check_site(mps, j)
    if i >= j:
        raise ValueError("string_order_parameter requires i < j.")
    ops = {i: left_operator, j: right_operator}
    for k in range(i + 1, j):
        ops[k] = string_operator
    return product_expectation(mps, ops, form=form)


def many_body_polarization(mps: Any, number_operator: Any, *, origin: Union[int, float] = 0, form: str = "B"
                           ) -> Dict[str, torch.Tensor]:
    """Resta polarisation: <U> with U = exp((2 pi i / L) sum_j (j + origin) n_j), a product of local unitaries
    (dmrg_advanced_observables.py:506-565).  Keys: polarization, expectation, phase."""
    n = torch.as_tensor(number_operator)
    N = mps.Nsites
    ops = {}
    for site in range(N):
        gen = (2j * torch.pi * float(site + origin) / N) * n.to(torch.complex128)
        ops[site] = torch.matrix_exp(gen)
    expectation = product_expectation(mps, ops, form=form)
    phase = torch.angle(expectation)
    return {"polarization": phase / (2.0 * torch.pi), "expectation": expectation, "phase": phase}


def _two_point_rows(ch: _Chain, M: torch.Tensor, rows) -> Tuple[torch.Tensor, torch.Tensor]:
    """(local values <O_i> for all i, matrix C[i, j] = <O_i O_j> for i in ``rows`` and all j), NOT normalised.
    For j > i the environment carrying O_i is pushed right one site at a time (one O-step to read C[i, j], one identity
    step to move on); for j < i the same from the right with the right-environment kernel."""
    N = ch.N
    local = torch.empty(N, device=ch.dev, dtype=ch.dt)
    for i in range(N):
        local[i] = ch.close(ch.step(ch.left(i), i, M), i)
    M2 = (M.reshape(ch.d, ch.d) @ M.reshape(ch.d, ch.d)).reshape(1, 1, ch.d, ch.d).contiguous()
    C = torch.zeros(len(rows), N, device=ch.dev, dtype=ch.dt)
    for r, i in enumerate(rows):
        C[r, i] = ch.close(ch.step(ch.left(i), i, M2), i)
        E = ch.step(ch.left(i), i, M)
        for j in range(i + 1, N):
            C[r, j] = ch.close(ch.step(E, j, M), j)
            if j + 1 < N:
                E = ch.step(E, j)
    return local, C


def structure_factor(mps: Any, operator: Any, q_values=None, *, connected: bool = False, form: str = "B"
                     ) -> Dict[str, torch.Tensor]:
    """S(q) = (1/L) sum_ij exp(i q (i - j)) C_ij with C_ij = <O_i O_j> (minus <O_i><O_j> if ``connected``)
    (dmrg_advanced_observables.py:184-261).  Operators on different sites commute, so C_ji = C_ij and only the upper
    triangle is contracted: N^2 / 2 pairs, two transfer steps each."""
    dt = _common_dtype(getattr(mps, form)[0].dtype if form in ("A", "B") else torch.float64, torch.as_tensor(operator).dtype)
    ch = _Chain(mps, form, dt)
    N = ch.N
    M = ch.op(operator)
    n2 = ch.norm2()
    if float(n2.abs()) == 0.0:
        raise ValueError("Cannot compute expectation value for a zero-norm state.")
    local, upper = _two_point_rows(ch, M, list(range(N)))
    local, upper = local / n2, upper / n2
    corr = torch.triu(upper) + torch.triu(upper, 1).T
    if not dt.is_complex:
        local, corr = local.real, corr.real
    if connected:
        corr = corr - local[:, None] * local[None, :]
    if q_values is None:
        q = 2.0 * torch.pi * torch.arange(N, device=ch.dev, dtype=torch.float64) / N
    else:
        q = torch.as_tensor(q_values, device=ch.dev, dtype=torch.float64).reshape(-1)
    pos = torch.arange(N, device=ch.dev, dtype=torch.float64)
    sep = pos[:, None] - pos[None, :]
    phase = torch.exp(1j * q[:, None, None] * sep[None])
    cdt = corr.dtype
    values = (phase.to(cdt) if cdt.is_complex else phase.real.to(cdt)) * corr[None]
    # the reference multiplies by phase.to(corr.dtype): for a real correlator only cos(q (i - j)) survives
    return {"q_values": q, "values": values.sum(dim=(1, 2)) / N, "correlation_matrix": corr}


def correlation_length(mps: Any, operator: Any, *, reference_site: Optional[int] = None,
                       fit_range: Optional[Tuple[int, int]] = None, connected: bool = True, form: str = "B",
                       eps: float = 1e-12) -> Dict[str, Any]:
    """xi from a least-squares fit of log|C(r)| against r for C(ref, j) (dmrg_advanced_observables.py:264-377)."""
    dt = _common_dtype(getattr(mps, form)[0].dtype if form in ("A", "B") else torch.float64, torch.as_tensor(operator).dtype)
    ch = _Chain(mps, form, dt)
    N = ch.N
    ref = N // 2 if reference_site is None else reference_site
    _check_site(mps, ref)
    M = ch.op(operator)
    n2 = ch.norm2()
    if float(n2.abs()) == 0.0:
        raise ValueError("Cannot compute expectation value for a zero-norm state.")
    local = torch.empty(N, device=ch.dev, dtype=ch.dt)
    for i in range(N):
        local[i] = ch.close(ch.step(ch.left(i), i, M), i)
    row = torch.zeros(N, device=ch.dev, dtype=ch.dt)
    E = ch.step(ch.left(ref), ref, M)                     # sites to the right of the reference
    for j in range(ref + 1, N):
        row[j] = ch.close(ch.step(E, j, M), j)
        if j + 1 < N:
            E = ch.step(E, j)
    F = ch.step_back(ch.right(ref), ref, M)               # sites to the left: the mirror image with right environments
    for j in range(ref - 1, -1, -1):
        row[j] = (ch.left(j) * ch.step_back(F, j, M)).sum()
        if j > 0:
            F = ch.step_back(F, j)
    local, row = local / n2, row / n2
    if connected:
        row = row - local[ref] * local
    sites = [j for j in range(N) if j != ref]
    rdt = torch.float64
    dist_t = torch.tensor([abs(j - ref) for j in sites], device=ch.dev, dtype=rdt)
    corr_t = row[sites] if dt.is_complex else row[sites].real
    if fit_range is None:
        r_min, r_max = 1, int(dist_t.max().item()) if dist_t.numel() else 0
    else:
        r_min, r_max = fit_range
        if r_min < 1 or r_max < r_min:
            raise ValueError("fit_range must be an inclusive window with 1 <= r_min <= r_max.")
    mag = corr_t.abs().to(rdt)
    mask = (dist_t >= r_min) & (dist_t <= r_max) & (mag > eps)
    if int(mask.sum()) < 2:
        raise ValueError("At least two nonzero correlation values are required to fit a correlation length.")
    X = torch.stack([dist_t[mask], torch.ones_like(dist_t[mask])], dim=1)
    sol = torch.linalg.lstsq(X.cpu(), torch.log(mag[mask]).cpu().unsqueeze(1)).solution.squeeze(1)
    slope, intercept = float(sol[0]), float(sol[1])
    if slope >= 0:
        raise ValueError("Fitted correlation decay slope is non-negative; cannot estimate a finite xi.")
    return {"xi": -1.0 / slope, "distances": dist_t, "correlations": corr_t, "fit_range": (int(r_min), int(r_max)),
            "slope": slope, "intercept": intercept}


def _padded(t: torch.Tensor, l: int, r: int, dt: torch.dtype) -> torch.Tensor:
    out = torch.zeros(l, t.shape[1], r, device=t.device, dtype=dt)
    out[: t.shape[0], :, : t.shape[2]] = t
    return out


def ground_state_fidelity(mps_a: Any, mps_b: Any, *, form_a: str = "B", form_b: str = "B") -> torch.Tensor:
    """|<a|b>| / (|a| |b|)  (dmrg_advanced_observables.py:471-503) by mixed transfer steps: the bra tensors of ``a``
    and the ket tensors of ``b`` (zero-padded to common bond dimensions) go through the row-block form of the
    left-environment kernel (``site_bra``)."""
    if mps_a.Nsites != mps_b.Nsites or mps_a.chid != mps_b.chid:
        raise ValueError("MPS objects must have matching Nsites and chid values.")
    from . import _lib
    Ta, Tb = _site_tensors(mps_a, form_a), _site_tensors(mps_b, form_b)
    dev = Ta[0].device
    if dev.type != "cuda":
        raise RuntimeError("qsb200 observables run on CUDA tensors only (no CPU fallback)")
    dt = _common_dtype(Ta[0].dtype, Tb[0].dtype)
    d = mps_a.chid
    eye = torch.eye(d, device=dev, dtype=dt).reshape(1, 1, d, d).contiguous()

    def overlap(X, Y):
        chi = max(X[0].shape[0], Y[0].shape[0])
        E = torch.eye(chi, device=dev, dtype=dt).reshape(1, chi, chi).contiguous()
        for x, y in zip(X, Y):
            l, r = max(x.shape[0], y.shape[0]), max(x.shape[2], y.shape[2])
            xb, yk = _padded(x.to(device=dev), l, r, dt), _padded(y.to(device=dev), l, r, dt)
            out = torch.empty(1, r, r, device=dev, dtype=dt)
            _lib.ops.env_left_rows(E, eye, yk, xb, out)
            E = out
        return torch.diagonal(E[0]).sum()
    na, nb = overlap(Ta, Ta).real, overlap(Tb, Tb).real
    if float(na) == 0.0 or float(nb) == 0.0:
        raise ValueError("Cannot compute fidelity for a zero-norm state.")
    return overlap(Ta, Tb).abs() / torch.sqrt(na * nb)


def _schmidt_probabilities(mps: Any, bond: int) -> torch.Tensor:
    if not isinstance(bond, int) or bond < 0 or bond >= len(mps.sWeight):
        raise ValueError(f"bond index {bond!r} is out of range for {len(mps.sWeight)} stored Schmidt vectors.")
    w = mps.sWeight[bond].detach()
    lam = torch.diagonal(w, 0) if w.dim() == 2 else w.reshape(-1)       # boundary weights are small matrices (:149-167)
    return (lam.abs() ** 2).real


def entanglement_spectrum(mps: Any, bond: Optional[int] = None, eps: float = 1e-15):
    """-log(p_alpha) of the normalised Schmidt probabilities above ``eps``, ascending
    (dmrg_advanced_observables.py:380-414); all stored bonds when ``bond`` is None."""
    def one(b):
        p = _schmidt_probabilities(mps, b)
        tot = p.sum()
        if float(tot) <= 0:
            return torch.empty(0, device=p.device, dtype=p.dtype)
        p = p / tot
        p = p[p > eps]
        return torch.sort(-torch.log(p)).values
    return [one(b) for b in range(len(mps.sWeight))] if bond is None else one(bond)


def entanglement_entropy(mps: Any, bond: Optional[int] = None, eps: float = 1e-15):
    """Von Neumann entropy -sum p log p of a bond's Schmidt probabilities (dmrg_observables.py:170-206)."""
    def one(b):
        p = _schmidt_probabilities(mps, b)
        tot = p.sum()
        if float(tot) <= 0:
            return 0.0
        p = p / tot
        p = p[p > eps]
        return float(-(p * torch.log(p)).sum())
    return [one(b) for b in range(len(mps.sWeight))] if bond is None else one(bond)


# ------------------------------------------------------------------------------------------------------------------
# host-side bookkeeping helpers of the reference's observables module (dmrg_observables.py:27-117, 209-233): no device
# work, kept so that code written against that module runs unchanged
# ------------------------------------------------------------------------------------------------------------------
def _as_float(value: Any) -> float:
    if isinstance(value, torch.Tensor):
        if value.numel() != 1:
            raise ValueError("Expected a scalar tensor.")
        return float(value.detach().real.cpu().item())
    return float(value.real) if isinstance(value, complex) else float(value)


def ground_state_energy(energies: Any) -> float:
    """Last entry of the energy trace ``DMRG.forward`` returns (dmrg_observables.py:51-76)."""
    seq = energies.reshape(-1) if isinstance(energies, torch.Tensor) else energies
    if len(seq) == 0:
        raise ValueError("Energy trace is empty.")
    return _as_float(seq[-1])


def energy_density(energy: Any, num_sites: int) -> float:
    """``energy / num_sites`` (dmrg_observables.py:79-94)."""
    if num_sites <= 0:
        raise ValueError("num_sites must be positive.")
    return _as_float(energy) / int(num_sites)


def excitation_gap(*args: Any, **kwargs: Any) -> float:
    """Reserved by the reference for a future excited-state solver (dmrg_observables.py:97-117); same behaviour."""
    raise NotImplementedError("Excitation gap requires excited-state DMRG or another E1 solver, "
                              "which is not implemented yet.")


def truncation_errors(dmrg_or_history: Any) -> List[float]:
    """Discarded weights recorded after every two-site split (dmrg_observables.py:209-233)."""
    from collections.abc import Mapping
    history = dmrg_or_history.history if hasattr(dmrg_or_history, "history") else dmrg_or_history
    if not isinstance(history, Mapping):
        raise TypeError("Expected a DMRG object with .history or a history mapping.")
    if "discarded_weights" not in history:
        raise KeyError("history does not contain 'discarded_weights'.")
    return [_as_float(v) for v in history["discarded_weights"]]

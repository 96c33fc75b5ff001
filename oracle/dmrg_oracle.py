"""CPU oracle for the two-site DMRG inner loop  --  TEST INFRASTRUCTURE ONLY.

This file restates, with plain ``torch`` CPU ops, the algorithm that
ToelUl/quantum-simulation executes on the hot path named in BASELINE.json:

  * H_eff matvec          reference ``quantum_simulation/dmrg.py:178-228`` driven
                          through ``ncon_torch.py:111-154`` (pairwise tensordot)
  * left / right env      ``dmrg.py:254-278`` / ``dmrg.py:284-308``
  * environment cache     ``dmrg.py:311-418``
  * restarted Lanczos     ``eigensolver/lanczos_solver.py:137-303``
  * random MPS            ``mps.py:21-70, 154-168``
  * two-site sweep driver ``dmrg.py:577-802``

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline.  The product package never imports it and
raises if its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` is importable) generated the fixtures in
``tests/golden/*.npz`` from the unmodified reference; ``tests/test_oracle.py``
checks every function here against them.

The arithmetic is the dense ncon order: each
pairwise step is one ``torch.tensordot`` exactly as the reference issues it,
so on the same MKL the results agree with the reference bit for bit.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def _pair(A: torch.Tensor, B: torch.Tensor, a_axes, b_axes) -> torch.Tensor:
    """One ncon pairwise step: tensordot over the common labels, the axis list
    ordered by the smaller operand's axis positions (ncon_torch.py:129-138)."""
    if A.numel() < B.numel():
        order = sorted(range(len(a_axes)), key=lambda k: a_axes[k])
    else:
        order = sorted(range(len(b_axes)), key=lambda k: b_axes[k])
    return torch.tensordot(A, B, dims=([a_axes[k] for k in order], [b_axes[k] for k in order]))


# --------------------------------------------------------------------------
# H_eff matvec (dmrg.py:216-228 via ncon_torch.py:111-154)
# --------------------------------------------------------------------------
def apply_mpo(psi_flat: torch.Tensor, L: torch.Tensor, M1: torch.Tensor,
              M2: torch.Tensor, R: torch.Tensor,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """H_eff . theta in the ncon order.

    Network (dmrg.py:218-221): theta(1,3,5,7) L(2,-1,1) M1(2,4,-2,3)
    M2(4,6,-3,5) R(6,-4,7); positive labels contracted in ascending order
    (ncon_torch.py:78-80).  The four steps below are the tensordots that order
    produces.
    """
    theta = psi_flat.view(L.shape[2], M1.shape[3], M2.shape[3], R.shape[2])
    # label 1: theta[a,s1,s2,b] x L[w,a',a]      -> (s1, s2, b, w, a')
    t1 = _pair(theta, L, [0], [2])
    # labels 2,3: M1[w,wm,s1',s1] x t1           -> (wm, s1', s2, b, a')
    t2 = _pair(M1, t1, [0, 3], [3, 0])
    # labels 4,5: M2[wm,wr,s2',s2] x t2          -> (wr, s2', s1', b, a')
    t3 = _pair(M2, t2, [0, 3], [0, 2])
    # labels 6,7: R[wr,b',b] x t3                -> (b', s2', s1', a')
    t4 = _pair(R, t3, [0, 2], [0, 3])
    res = t4.permute(3, 2, 1, 0)                 # (a', s1', s2', b')  (ncon_torch.py:179-183)
    if out is not None:
        h, i, j, k = res.shape
        if out.numel() != h * i * j * k or out.device != res.device or out.dtype != res.dtype:
            raise ValueError("[ApplyMPO] out shape/device/dtype mismatch.")
        out.view(h, i, j, k).copy_(res)          # dmrg.py:226
        return out.view(-1)
    return res.reshape(-1)                       # dmrg.py:228


def apply_mpo_einsum(psi_flat, L, M1, M2, R):
    """The reference's other branch (dmrg.py:230-248); used as a cross-check."""
    b, h, a = L.shape
    _b, d, i, c = M1.shape
    _d, f, j, e = M2.shape
    _f, k, g = R.shape
    assert _b == b and _d == d and _f == f
    psi = psi_flat.view(a, c, e, g)
    t01 = torch.einsum('aceg,bha->bcegh', psi, L)
    t02 = torch.einsum('bcegh,fkg->bcehfk', t01, R)
    t03 = torch.einsum('bcehfk,dfje->bcdhjk', t02, M2)
    res = torch.einsum('bcdhjk,bdic->hijk', t03, M1)
    return res.contiguous().view(-1)


# --------------------------------------------------------------------------
# environment updates (dmrg.py:254-308)
# --------------------------------------------------------------------------
def left_env_update(Lp: torch.Tensor, M: torch.Tensor, Ap: torch.Tensor) -> torch.Tensor:
    """L[p+1] from L[p], M[p], A[p].

    Network (dmrg.py:275-278): Lp(2,1,4) M(2,-1,3,5) A(4,5,-3) conjA(1,3,-2).
    Returned, like the reference, as the permuted view (W', chi'_bra, chi'_ket).
    """
    cA = torch.conj(Ap)
    t1 = _pair(Lp, cA, [1], [0])           # (w, a_ket, s', c'_bra)
    t2 = _pair(M, t1, [0, 2], [0, 2])      # (w', s, a_ket, c'_bra)
    t3 = _pair(Ap, t2, [0, 1], [2, 1])     # (c_ket, w', c'_bra)
    return t3.permute(1, 2, 0)                              # (w', c'_bra, c_ket)


def right_env_update(M: torch.Tensor, Rnext: torch.Tensor, Bp1: torch.Tensor) -> torch.Tensor:
    """R[p] from M[p], R[p+1], B[p].

    Network (dmrg.py:305-308): M(-1,2,3,5) Rnext(2,1,4) B(-3,5,4) conjB(-2,3,1).
    """
    cB = torch.conj(Bp1)
    t1 = _pair(Rnext, cB, [1], [2])        # (w', b_ket, a'_bra, s')
    t2 = _pair(M, t1, [1, 2], [0, 3])      # (w, s, b_ket, a'_bra)
    t3 = _pair(Bp1, t2, [2, 1], [2, 1])    # (a_ket, w, a'_bra)
    return t3.permute(1, 2, 0)                              # (w, a'_bra, a_ket)


class EnvCache:
    """Restatement of EnvironmentManager (dmrg.py:311-418)."""

    def __init__(self, mpo: Sequence[torch.Tensor], dtype: torch.dtype):
        self.mpo = list(mpo)
        self.N = len(self.mpo)
        self.L: List[Optional[torch.Tensor]] = [None] * (self.N + 1)
        self.R: List[Optional[torch.Tensor]] = [None] * (self.N + 1)
        self.L[0] = torch.ones(self.mpo[0].shape[0], 1, 1, dtype=dtype)          # dmrg.py:348-349
        self.R[self.N] = torch.ones(self.mpo[-1].shape[1], 1, 1, dtype=dtype)    # dmrg.py:350-351

    def get_L(self, p: int) -> torch.Tensor:
        if self.L[p] is None:
            raise RuntimeError(f"L[{p}] requested but not yet computed.")
        return self.L[p]

    def get_R(self, p: int) -> torch.Tensor:
        if self.R[p] is None:
            raise RuntimeError(f"R[{p}] requested but not yet computed.")
        return self.R[p]

    def update_L(self, p: int, A: torch.Tensor) -> None:
        self.L[p + 1] = left_env_update(self.get_L(p), self.mpo[p], A)

    def update_R(self, p: int, B: torch.Tensor) -> None:
        self.R[p] = right_env_update(self.mpo[p], self.get_R(p + 1), B)


# --------------------------------------------------------------------------
# restarted Lanczos (lanczos_solver.py:137-303)
# --------------------------------------------------------------------------
def _real_dtype(dtype: torch.dtype) -> torch.dtype:
    return {torch.complex128: torch.float64, torch.complex64: torch.float32}.get(dtype, dtype)


def _normalize_to(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    out.copy_(x)                                             # lanczos_solver.py:148
    nrm = torch.linalg.norm(out)
    if nrm == 0:                                             # :150-153 zero vector re-seed
        out.uniform_(-1, 1)
        nrm = torch.linalg.norm(out)
    out.div_(nrm)
    return out


def lanczos_ground_state(initial_vector: torch.Tensor, linear_operator: Callable,
                         operator_args: Tuple = (), num_restarts: int = 2,
                         krylov_dim: int = 4, *, pro: bool = False,
                         pro_max_reorth: int = 1, beta_tol: float = 1e-12,
                         ritz_residual_tol: Optional[float] = None
                         ) -> Tuple[torch.Tensor, float]:
    """Same control flow and the same torch ops, op for op, as _LanczosEngine.forward."""
    with torch.inference_mode():
        dtype = initial_vector.dtype
        rdt = _real_dtype(dtype)
        n = initial_vector.numel()
        k = max(1, int(krylov_dim))
        K = torch.empty(n, k, dtype=dtype)
        alphas = torch.empty(k, dtype=rdt)
        betas = torch.empty(max(0, k - 1), dtype=rdt)
        T = torch.empty(k, k, dtype=rdt)
        vec_cur = torch.empty(n, dtype=dtype)
        vec_prev = torch.empty(n, dtype=dtype)
        w = torch.empty(n, dtype=dtype)
        tmp_n = torch.empty(n, dtype=dtype)
        tmp_n2 = torch.empty(n, dtype=dtype)
        ov = torch.empty(k, dtype=dtype)

        best_vec = _normalize_to(initial_vector, vec_cur.clone())
        best_e = torch.tensor(float("inf"), dtype=rdt)

        for _ in range(num_restarts):
            _normalize_to(best_vec, vec_cur)
            alphas.zero_()
            if betas.numel() > 0:
                betas.zero_()
            beta_prev = None
            beta_tail = torch.tensor(0.0, dtype=rdt)
            c = 0
            for j in range(k):
                K[:, j].copy_(vec_cur)
                c += 1
                try:
                    linear_operator(vec_cur, *operator_args, out=w)
                except TypeError:
                    w.copy_(linear_operator(vec_cur, *operator_args))
                alpha_j = torch.vdot(vec_cur, w).real.to(rdt)
                alphas[j] = alpha_j
                w.add_(vec_cur, alpha=-alpha_j.to(dtype))
                if beta_prev is not None:
                    w.add_(vec_prev, alpha=-beta_prev.to(dtype))
                if pro and j > 0:
                    Kc = K[:, :j]
                    ovj = ov[:j]
                    for _r in range(pro_max_reorth):
                        torch.addmv(ovj, Kc.conj().transpose(0, 1), w, beta=0.0, alpha=1.0, out=ovj)
                        torch.addmv(tmp_n, Kc, ovj, beta=0.0, alpha=1.0, out=tmp_n)
                        w.add_(tmp_n, alpha=-1.0)
                beta_j = torch.linalg.norm(w).to(rdt)
                last = (j == k - 1)
                if beta_j.item() < beta_tol or last:
                    beta_tail = beta_j
                    break
                vec_prev.copy_(vec_cur)
                vec_cur.copy_(w)
                vec_cur.div_(beta_j.to(dtype))
                betas[j] = beta_j
                beta_prev = beta_j

            Tloc = T[:c, :c]
            Tloc.zero_()
            Tloc.diagonal().copy_(alphas[:c])
            if c > 1:
                Tloc.diagonal(1).copy_(betas[:c - 1])
                Tloc.diagonal(-1).copy_(betas[:c - 1])
            evals, evecs = torch.linalg.eigh(Tloc)
            order = torch.argsort(evals.real)
            theta = evals[order[0]].real.to(rdt)
            y = evecs[:, order[0]]
            if theta < best_e:
                torch.addmv(tmp_n, K[:, :c], y.to(dtype), beta=0.0, alpha=1.0, out=tmp_n)
                best_vec = _normalize_to(tmp_n, tmp_n2)
                best_e = theta
            if ritz_residual_tol is not None:
                if (beta_tail.abs() * y[-1].abs()).item() < ritz_residual_tol:
                    break
        return best_vec, float(best_e.item())


# --------------------------------------------------------------------------
# random right-canonical MPS (mps.py:21-70, 154-168)
# --------------------------------------------------------------------------
class OracleMPS:
    """Plain-list restatement of MPS: A / B / sWeight lists, ortho_center = 0."""

    def __init__(self, Nsites: int, chid: int, chi: int, dtype: torch.dtype, seed: Optional[int] = None):
        self.Nsites, self.chid, self.dtype = Nsites, chid, dtype
        self.ortho_center = 0
        if seed is not None:
            g = torch.Generator(device="cpu").manual_seed(seed)
            rand = lambda *sh: torch.rand(*sh, generator=g, dtype=dtype)
        else:
            rand = lambda *sh: torch.rand(*sh, dtype=dtype)
        A = [None] * Nsites
        A[0] = rand(1, chid, min(chi, chid))                                   # mps.py:163
        for s in range(1, Nsites):
            left = A[s - 1].shape[2]
            right = min(min(chi, left * chid), chid ** (Nsites - s - 1))       # mps.py:166
            A[s] = rand(left, chid, right)
        for p in range(Nsites - 1, 0, -1):                                     # mps.py:52-58
            chil, d, chir = A[p].shape
            Q, Rm = torch.linalg.qr(A[p].reshape(chil, d * chir).T)
            A[p] = Q.T.reshape(Q.shape[1], d, chir)
            A[p - 1] = torch.einsum('lsc,ck->lsk', A[p - 1], Rm.T)
        A[0] = A[0] / torch.linalg.norm(A[0])
        self.A = A
        self.B = [t.clone() for t in A]
        self.sWeight = [torch.ones(1, dtype=dtype) for _ in range(Nsites + 1)]


# --------------------------------------------------------------------------
# two-site DMRG driver (dmrg.py:577-802), noise-free
# --------------------------------------------------------------------------
def dmrg_run(mps: OracleMPS, mpo: Sequence[torch.Tensor], *, numsweeps: int,
             maxdim: Sequence[int], krylov_dim: Sequence[int], cutoff: Sequence[float],
             reortho: Sequence[bool], maxreortho: Sequence[int], maxit: int = 2,
             matvec: Callable = apply_mpo):
    """Returns (Ekeep, history) with the reference's history keys (dmrg.py:566-575)."""
    dtype = mps.dtype
    Ms = [m.to(dtype) for m in mpo]
    N, d = len(Ms), Ms[0].shape[2]
    env = EnvCache(Ms, dtype)
    for p in range(N - 1, mps.ortho_center, -1):                               # dmrg.py:752-753
        env.update_R(p, mps.B[p])
    Ekeep: List[float] = []
    hist = {k: [] for k in ('energies', 'discarded_weights', 'bond_dims', 'sweeps', 'directions', 'sites')}
    A, B, sW = mps.A, mps.B, mps.sWeight
    eps = torch.finfo(_real_dtype(dtype)).eps

    def half(direction: str, kk: int):
        chi, kd, cut = maxdim[kk], krylov_dim[kk], cutoff[kk]
        rng = range(N - 1) if direction == 'right' else range(N - 2, -1, -1)
        for p in rng:
            if direction == 'right':                                          # dmrg.py:633-646
                chil, chir = B[p].shape[0], B[p + 1].shape[2]
                sw = sW[p]
                if sw.dim() == 2:
                    sw = torch.diagonal(sw, 0)
                sw = sw.to(B[p].dtype)
                l, s, m = B[p].shape
                _m, t, r = B[p + 1].shape
                psi = torch.matmul(torch.mul(B[p], sw.view(-1, 1, 1)).reshape(l * s, m),
                                   B[p + 1].reshape(m, t * r))
            else:                                                             # dmrg.py:647-660
                chil, chir = A[p].shape[0], A[p + 1].shape[2]
                sw = sW[p + 2]
                if sw.dim() == 2:
                    sw = torch.diagonal(sw, 0)
                sw = sw.to(A[p + 1].dtype)
                l, s, m = A[p].shape
                _m, t, r = A[p + 1].shape
                psi = torch.matmul(A[p].reshape(l * s, m),
                                   torch.mul(A[p + 1], sw.view(1, 1, -1)).reshape(m, t * r))
            flat = psi.view(-1)
            flat, E = lanczos_ground_state(flat, matvec, (env.get_L(p), Ms[p], Ms[p + 1], env.get_R(p + 2)),
                                           num_restarts=maxit, krylov_dim=kd,
                                           pro=reortho[kk], pro_max_reorth=maxreortho[kk])
            Ekeep.append(E)
            hist['energies'].append(float(E))
            U, S, Vh = torch.linalg.svd(flat.view(chil * d, d * chir), full_matrices=False)   # dmrg.py:685
            keep_max = min(S.numel(), chi)
            keep_cut = S.numel()
            if cut > 0.0:                                                     # dmrg.py:689-692
                S2 = S ** 2
                cs = torch.cumsum(S2 / torch.sum(S2), dim=0)
                keep_cut = int(torch.searchsorted(cs, 1.0 - cut, right=True)) + 1
            keep = max(1, min(keep_max, keep_cut))
            hist['discarded_weights'].append(torch.sum(S[keep:] ** 2).real.item())
            hist['bond_dims'].append(int(keep))
            hist['sweeps'].append(kk + 1)
            hist['directions'].append(direction)
            hist['sites'].append(int(p))
            sv = S[:keep]
            denom = sv.norm().add_(torch.finfo(sv.dtype).eps)
            A[p] = U[:, :keep].view(chil, d, keep)
            sW[p + 1] = sv.div(denom)
            B[p + 1] = Vh[:keep, :].view(keep, d, chir)
            if direction == 'right':
                env.update_L(p, A[p])
            else:
                env.update_R(p + 1, B[p + 1])

    for kk in range(numsweeps):
        half('right', kk)
        mps.ortho_center = N - 1
        sw = sW[N - 1]                                                         # dmrg.py:770-781
        if sw.dim() == 2:
            sw = torch.diagonal(sw, 0)
        Bl = B[N - 1]
        At = (sw.to(Bl.dtype).view(-1, 1, 1) * Bl).reshape(Bl.shape[0] * d, Bl.shape[2])
        U, S, Vh = torch.linalg.svd(At, full_matrices=False)
        A[N - 1] = U.view(Bl.shape[0], d, Bl.shape[2])
        sW[N] = (S.unsqueeze(1) * Vh).to(dtype) / (S.norm().add_(eps))
        half('left', kk)
        mps.ortho_center = 0
        sw = sW[1]                                                             # dmrg.py:788-796
        if sw.dim() == 2:
            sw = torch.diagonal(sw, 0)
        A0 = A[0] * sw.to(A[0].dtype).view(1, 1, -1)
        chil, _, chir = A[0].shape
        U, S, Vh = torch.linalg.svd(A0.reshape(chil, d * chir), full_matrices=False)
        B[0] = Vh.view(chil, d, chir)
        sW[0] = (U * S.unsqueeze(0)).to(dtype) / (S.norm().add_(eps))
    return Ekeep, hist


# --------------------------------------------------------------------------
# algorithmic flop counts
# --------------------------------------------------------------------------
def matvec_flops(L, M1, M2, R) -> float:
    """Dense ncon-order flop count of one H_eff matvec; c = 2 real, 8 complex."""
    c = 8.0 if L.dtype.is_complex else 2.0
    Wl, clo, cli = L.shape
    _, Wm, d1o, d1i = M1.shape
    _, Wr, d2o, d2i = M2.shape
    _, cro, cri = R.shape
    return c * (Wl * clo * cli * d1i * d2i * cri
                + (Wm * d1o) * (Wl * d1i) * (d2i * cri * clo)
                + (Wr * d2o) * (Wm * d2i) * (d1o * cri * clo)
                + cro * (Wr * cri) * (d2o * d1o * clo))


def env_flops(W: int, Wp: int, d: int, chi: int, chip: int, is_complex: bool) -> float:
    c = 8.0 if is_complex else 2.0
    return c * (W * d * chi * chi * chip + W * Wp * d * d * chi * chip + Wp * d * chi * chip * chip)

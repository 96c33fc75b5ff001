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

Truncating blocks (``maxdim + oversample <= 0.65 min(m, n)``: every bulk bond at a saturated bond dimension,
where theta is 2 chi x 2 chi and chi directions are kept) first go through a RANGE COMPRESSION (``_sketch_split``):
``Y = theta Omega`` with a fixed Gaussian ``Omega`` (n x l, l = maxdim + min(256, maxdim/8)), ``Q = orth(Y)`` by shifted CholeskyQR3
on the GEMM core, ``B = Q^H theta`` (l x n), and then the same Gram/eigh factorisation of the SMALL matrix B (eigh of
l x l instead of min(m,n) x min(m,n): 14 ms instead of 85 ms at chi = 2048).  This is exact up to the weight theta has
outside range(Q), ``rho = |theta|^2 - |B|^2``, which bounds the error of every squared Schmidt value
(``lambda_j(B B^H) <= lambda_j(theta theta^H) <= lambda_j(B B^H) + rho``).  The result is accepted only if
``rho <= SKETCH_TOL |theta|^2`` (1e-13: the absolute noise floor the Gram method has anyway, eps * lambda_max summed
over the tail); otherwise the block is redone by the full Gram/eigh path, and the sketch is skipped for the next few
blocks of that shape.  On row shards (and optionally on one GPU, ``SKETCH_REFINE_SINGLE``) a
marginally rejected compression (rho <= 2e-11) first gets one step of subspace iteration with re-orthonormalisation
(V = orth(B^H), Y <- theta V), which sharpens rho from ~50 x to ~1.6 x the weight beyond the compressed dimension.  (The plain power step theta theta^H Q was tried and dropped: it squares the spectrum and
pushes the directions that matter below the regularisation level.)  ``S_all`` then carries the l leading Schmidt values plus one pseudo-value sqrt(rho), so that
``sum(S_all[keep:]**2)`` is still the discarded weight the caller records.
"""
from __future__ import annotations

from typing import Tuple

import torch

__all__ = ["two_site_split", "sketch_split_sharded", "gram_split_sharded", "sharded_split_eligible", "FAST_MIN"]

FAST_MIN = 256        # min(m, n) from which the Gram/eigh path is used
SKETCH = True         # range compression in front of the Gram/eigh path for truncating blocks (module docstring)
SKETCH_OVERSAMPLE = 256
SKETCH_MAX_FRACTION = 0.65     # use it when maxdim + oversample <= this fraction of min(m, n)
SKETCH_TOL = 1e-13             # accepted weight of theta outside the compressed range, relative to |theta|^2
SKETCH_REG = 1e-9              # relative size of the regularising perturbation of the sketch (see _sketch_split)
SKETCH_REFINE_BELOW = 2e-11    # a rejected compression with rho below this gets one subspace-iteration refinement ...
SKETCH_REFINE_SINGLE = False   # ... on row shards always (the gathered rank-0 fallback costs 3x more); on one GPU it
                               # measured as a wash against falling back (C3 sweep 36.4 -> 38.7 s with it), so it is off
SKETCH_ADAPT = True            # rank-adaptive compression width (see _sketch_split): narrow sketch for over-resolved blocks
SKETCH_TOL_NARROW = 1e-15      # accepted residual weight of a NARROW compression (the rest of the kept basis is padding)
SKETCH_RANK_WEIGHT = 1e-17     # directions whose tail weight is below this do not count towards the numerical rank
SKETCH_BACKOFF = 4             # after a rejected sketch: blocks of the same shape that skip it
stats = {"sketch_accepted": 0, "sketch_rejected": 0, "sketch_skipped": 0, "gram": 0, "svd": 0, "last_rho": None}
_omega_cache: dict = {}
_backoff: dict = {}
_rank_hint: dict = {}          # (m, n, k) -> numerical rank seen by the last accepted compression of that shape


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


def _orthonormal_rows(B: torch.Tensor, allsum=None, agree=None):
    """Rows of B (k x n, k <= n, mutually orthogonal up to rounding, norms spanning many decades) -> orthonormal rows
    spanning the same flag of subspaces (row i stays a combination of rows <= i, positive on itself: the gauge of a
    QR factorisation of B^H with a real positive diagonal).  CholeskyQR2 on the row-normalised matrix; None if a
    Cholesky factorisation breaks down (the caller then uses Householder QR).
    Column-sharded use (sketch_split_sharded): ``B`` is this rank's column block, ``allsum(t)`` sums a tensor over the
    ranks in place and ``agree(flag)`` returns the OR of a failure flag over the ranks (one decision for everyone)."""
    nrm2 = torch.linalg.norm(B, dim=1) ** 2
    if allsum is not None:
        allsum(nrm2)
    nrm = torch.sqrt(nrm2)
    bad = not bool((nrm > 0).all())
    if agree(bad) if agree is not None else bad:
        return None
    X = B / nrm.unsqueeze(1).to(B.dtype)
    for _ in range(2):
        Gm = _mm(X, X.T, conj_b=True)                        # k x k, = I + small
        if allsum is not None:
            allsum(Gm)
        Gm = 0.5 * (Gm + Gm.conj().T)
        Lc, info = torch.linalg.cholesky_ex(Gm)
        bad = int(info) != 0
        if agree(bad) if agree is not None else bad:
            return None
        X = torch.linalg.solve_triangular(Lc, X, upper=False)     # X <- Lc^{-1} X : rows orthonormal
    return X


def _orthonormal_cols(Y: torch.Tensor, max_passes: int = 6, *, m_total: int = None, allsum=None, agree=None):
    """Orthonormal basis of range(Y) for a tall, ill-conditioned Y (m x l) by adaptive shifted CholeskyQR (Fukaya et
    al.): every pass forms the Gram matrix G = X^H X (GEMM on the library's core), factors it and solves X <- X L^-H.
    The first pass -- and any later pass whose plain Cholesky factorisation breaks down -- adds the shift
    s = 11 (m l + l (l + 1)) eps trace(G), which divides the condition number by ~1e4 per pass; as soon as G is within
    1e-2 of the identity one last plain pass makes X orthonormal to O(eps).  Typically 1 shifted + 2 plain passes at the
    regularisation used by the callers, one more shifted pass for harder spectra.  None if ``max_passes`` do not get
    there (the caller then uses Householder QR).
    Row-sharded use: ``Y`` is this rank's row block of an ``m_total`` x l matrix; ``allsum`` / ``agree`` as in
    ``_orthonormal_rows`` (the l x l Gram matrices are summed over the ranks, the Cholesky factorisations replicated,
    every branch is decided from all-reduced data)."""
    m, l = Y.shape
    m = m if m_total is None else int(m_total)
    eps = torch.finfo(Y.real.dtype if Y.is_complex() else Y.dtype).eps
    eye = None
    X = Y
    shifted = True
    it = 0
    while it < max_passes:
        Gm = _mm(X.T, X, conj_a=True)                         # l x l
        if allsum is not None:
            allsum(Gm)
        Gm = 0.5 * (Gm + Gm.conj().T)
        if it > 0:
            if eye is None:
                eye = torch.eye(l, dtype=Gm.dtype, device=Gm.device)
            if float((Gm - eye).abs().max()) < 1e-2:          # last pass: plain Cholesky of a matrix close to I
                Lc, info = torch.linalg.cholesky_ex(Gm)
                bad = int(info) != 0
                if agree(bad) if agree is not None else bad:
                    return None
                return torch.linalg.solve_triangular(Lc.conj().T, X, upper=True, left=False)
        G2 = Gm
        if shifted:
            G2 = Gm.clone()
            G2.diagonal().add_((11.0 * (m * l + l * (l + 1)) * eps * torch.diagonal(Gm).real.sum()).to(Gm.dtype))
        Lc, info = torch.linalg.cholesky_ex(G2)
        bad = int(info) != 0
        if agree(bad) if agree is not None else bad:
            if shifted:
                return None                                    # not even the shifted matrix is positive definite
            shifted = True                                     # redo this pass with the shift (X unchanged)
            it += 1
            continue
        X = torch.linalg.solve_triangular(Lc.conj().T, X, upper=True, left=False)     # X <- X Lc^{-H}
        shifted = False
        it += 1
    return None


def clear_caches() -> None:
    """Free the cached test / regularisation matrices of the range compression (up to 8 matrices of n x l elements)
    and forget the per-shape back-off state."""
    _omega_cache.clear()
    _backoff.clear()
    _backoff_dist.clear()
    _rank_hint.clear()


def _omega(n: int, l: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """The fixed test matrix of the range compression: standard normal, generated once per (n, l, dtype, device) from
    its own seeded generator (no global RNG state is consumed; results repeat run to run)."""
    return _fixed_normal("omega", n, l, dtype, device, 0x51EE7)


def _fixed_normal(tag: str, rows: int, cols: int, dtype: torch.dtype, device: torch.device, seed: int) -> torch.Tensor:
    key = (tag, rows, cols, dtype, str(device))
    t = _omega_cache.get(key)
    if t is None:
        if len(_omega_cache) >= 8:
            _omega_cache.clear()
        g = torch.Generator(device=device).manual_seed(seed)
        t = torch.randn(rows, cols, dtype=torch.float64, device=device, generator=g).to(dtype)
        _omega_cache[key] = t
    return t


def _complete_isometry(U1: torch.Tensor, c: int, tag: str) -> torch.Tensor:
    """c more orthonormal columns, orthogonal to the orthonormal columns of U1 (m x r): a fixed Gaussian block with U1
    projected out, orthonormalised (CholeskyQR; the projected Gaussian is well conditioned), U1 projected out once more
    (an O(eps) correction that leaves the columns orthonormal to O(eps)).  The padding of an over-resolved bond:
    directions that carry no weight, only there because the reference's MPS has exactly ``maxdim`` columns."""
    m = U1.shape[0]
    X = _fixed_normal(tag, m, c, U1.dtype, U1.device, 0xC0DE)
    X = X - _mm(U1, _mm(U1.T, X, conj_a=True))
    Q = _orthonormal_cols(X)
    X = Q if Q is not None else torch.linalg.qr(X)[0]
    return X - _mm(U1, _mm(U1.T, X, conj_a=True))


def _sketch_split(theta: torch.Tensor, maxdim: int, cutoff: float, _width: int = None):
    """Range compression + Gram/eigh of the compressed block (module docstring).  Returns None when the a-posteriori
    bound rejects the compression (the caller falls back to the full Gram/eigh path).

    Rank-adaptive width (``SKETCH_ADAPT``): the last accepted compression of a shape records the numerical rank r of
    the block (directions beyond it carry < 1e-17 of the weight -- far below what any Gram method resolves).  The next
    block of that shape is compressed to l1 = 1.5 r + 32 columns only; if theta has <= 1e-15 of its weight outside that
    range, the l1 x l1 eigenproblem replaces the (maxdim + p)^2 one, and, when l1 < maxdim, the kept basis is padded
    with weightless orthonormal directions (Schmidt values 0) up to the bond dimension the reference would keep.
    An over-resolved run (TFIM at chi = 512: rank ~30) then pays for its rank, not for its bond dimension; otherwise
    the narrow attempt is skipped (hint >= full width) or retried at full width."""
    m, n = theta.shape
    k = min(int(maxdim), m, n)
    l_full = k + _oversample(k)
    if m > n:                                                  # compress the smaller side's partner: work on theta^H
        out = _sketch_split(theta.conj().T, maxdim, cutoff)
        if out is None:
            return None
        U, S, Vh, keep = out
        return Vh.conj().T.resolve_conj().contiguous(), S, U.conj().T.resolve_conj().contiguous(), keep
    rdt = theta.real.dtype if theta.is_complex() else theta.dtype
    key = (m, n, k)
    l = l_full if _width is None else _width
    hint = _rank_hint.get(key)
    if _width is None and SKETCH_ADAPT and hint is not None:
        l = min(l_full, (max(64, int(1.5 * hint) + 32) + 63) // 64 * 64)
    narrow = l < l_full
    Y = _mm(theta, _omega(n, l_full, theta.dtype, theta.device)[:, :l])       # m x l
    # DMRG blocks are numerically rank-deficient (Schmidt values far below eps), which no Cholesky-based
    # orthogonalisation survives: `orth` lifts the null directions of Y to SKETCH_REG |Y| with a fixed Gaussian
    # perturbation.  range(Q) then contains the resolved directions of range(Y) up to angles ~SKETCH_REG / sigma_i, which
    # costs ~rank * SKETCH_REG^2 of captured weight (1e-19 measured at 1e-9) -- measured by rho like everything else.
    tot = torch.linalg.norm(theta) ** 2

    def orth(Yc, tag):
        """Regularised shifted CholeskyQR of a tall matrix (Householder QR when a Cholesky pivot fails)."""
        if SKETCH_REG > 0:
            Yc.add_(_fixed_normal(tag, Yc.shape[0], l_full, theta.dtype, theta.device, 0xB200)[:, :Yc.shape[1]],
                    alpha=float(SKETCH_REG * torch.linalg.norm(Yc)) / (Yc.shape[0] * Yc.shape[1]) ** 0.5)
        Qc = _orthonormal_cols(Yc)
        if Qc is None:
            stats["householder"] = stats.get("householder", 0) + 1
            Qc = torch.linalg.qr(Yc)[0]
        return Qc

    def compress(Yc):
        """Q = orth(Yc), B = Q^H theta and the weight of theta outside range(Q)."""
        Qc = orth(Yc, "reg")
        Bc = _mm(Qc.T, theta, conj_a=True)                     # l x n
        return Qc, Bc, torch.linalg.norm(theta - _mm(Qc, Bc)) ** 2      # (|theta|^2 - |B|^2 would cancel at 1e-14)
    Q, B, rho = compress(Y)
    stats["last_rho"] = float(rho / tot)
    if narrow and not stats["last_rho"] <= SKETCH_TOL_NARROW:  # the block outgrew the hint: full width
        stats["sketch_narrow_retry"] = stats.get("sketch_narrow_retry", 0) + 1
        _rank_hint.pop(key, None)
        return _sketch_split(theta, maxdim, cutoff, _width=l_full)
    if SKETCH_REFINE_SINGLE and SKETCH_TOL < stats["last_rho"] <= SKETCH_REFINE_BELOW:
        # marginal miss: one step of subspace iteration WITH re-orthonormalisation (V = orth(B^H), Y <- theta V): the
        # captured range sharpens from ~(1 + k/p) tail to ~tail at the cost of two more orthonormalisations -- still
        # cheaper than the full eigenproblem, much cheaper than the gathered fallback of the sharded driver
        stats["sketch_refined"] = stats.get("sketch_refined", 0) + 1
        V1 = orth(B.conj().T.resolve_conj().contiguous(), "reg2")      # n x l
        Q, B, rho = compress(_mm(theta, V1))
        stats["last_rho"] = float(rho / tot)
    if not stats["last_rho"] <= SKETCH_TOL:                    # (NaN-safe)
        stats.setdefault("rejected_rhos", []).append(stats["last_rho"])
        return None
    Gs = _mm(B, B.T, conj_b=True)                              # l x l
    Gs = 0.5 * (Gs + Gs.conj().T)
    lam, vecs = torch.linalg.eigh(Gs)
    lam = torch.flip(lam, dims=(0,)).clamp_min(0)
    tail = torch.cumsum(lam.flip(0), 0).flip(0) / tot          # weight carried from index j on
    _rank_hint[key] = int((tail > SKETCH_RANK_WEIGHT).sum())
    if TIMING:                                                 # diagnostics: how many directions carry weight above ...
        stats.setdefault("rank_tail_1e-15", []).append(int((tail > 1e-15).sum()))
        stats.setdefault("rank_tail_1e-13", []).append(int((tail > 1e-13).sum()))
    if narrow:
        stats["sketch_narrow"] = stats.get("sketch_narrow", 0) + 1
    S = torch.cat([torch.sqrt(lam), torch.sqrt(rho).reshape(1).to(rdt)])      # leading l values + the residual weight
    # the reference's truncation rule (dmrg.py:687-695) over the full spectrum: the total weight is known (|theta|^2)
    # and a cut beyond position l can only raise keep_cut above keep_max = k < l
    if l > k:
        keep = _keep_from_spectrum(S, k, cutoff)
        pad = 0
    else:                                                      # narrow compression: l <= k resolved directions
        cut = _keep_from_spectrum(S, 10 ** 9, cutoff) if cutoff > 0.0 else l + 1
        keep = cut if cut <= l else k                          # a cut beyond the resolved part keeps the full bond
        pad = max(0, keep - l)
    take = keep - pad
    top = torch.flip(vecs[:, vecs.shape[1] - take:], dims=(1,))                 # l x take
    U = _mm(Q, top)                                            # m x take, orthonormal columns
    Bk = _mm(top.T, B, conj_a=True)                            # take x n : S_k Vh_k
    Vh = _orthonormal_rows(Bk) if USE_CHOLQR else None
    if Vh is None:
        Qc, Rr = torch.linalg.qr(Bk.conj().T)
        Vh = _fix_phase(Qc, Rr).conj().T
    if pad:                                                    # weightless padding up to the reference's bond dimension
        U = torch.cat([U, _complete_isometry(U, pad, "padU")], dim=1)
        Vh = torch.cat([Vh, _complete_isometry(Vh.conj().T.resolve_conj().contiguous(), pad, "padV").conj().T], dim=0)
        S = torch.cat([S[:take], torch.zeros(pad, dtype=rdt, device=S.device), S[-1:]])
    return U.resolve_conj().contiguous(), S, Vh.resolve_conj().contiguous(), keep


_backoff_dist: dict = {}      # like _backoff, but only touched by collective calls (identical on every rank)


def sharded_split_eligible(m: int, n: int, maxdim: int, world: int) -> bool:
    """Can ``sketch_split_sharded`` take a block of this shape?  (A truncating block whose rows and columns both
    partition evenly; everything else is gathered and split on rank 0.)"""
    k = min(int(maxdim), m, n)
    return (SKETCH and world > 1 and m % world == 0 and n % world == 0 and min(m, n) >= FAST_MIN
            and k + _oversample(k) <= SKETCH_MAX_FRACTION * min(m, n))


def sketch_split_sharded(theta_rows: torch.Tensor, m: int, maxdim: int, cutoff: float, group, comm: dict = None):
    """The range-compressed two-site split (``_sketch_split``) on a ROW-SHARDED theta: rank r holds rows
    [r m/P, (r+1) m/P) of the m x n block (what the sharded Lanczos solver returns) and nobody gathers it.

      Y_r = theta_r Omega                       local GEMM
      Q   = orth(Y)  (shifted CholeskyQR)       l x l Gram matrices all-reduced, Cholesky replicated, solves local
      B   = sum_r Q_r^H theta_r                 local GEMM + all-reduce (l x n)
      rho = sum_r |theta_r - Q_r B|^2           acceptance test, one decision for all ranks
      G_s = sum_c B_c B_c^H                     column blocks, all-reduce (l x l)
      eigh(G_s)                                 rank 0, leading eigenvectors + spectrum broadcast (bit-identical S)
      U_r = Q_r W_k        -> all-gather        (m x keep), replicated like the rest of the MPS
      Vh_c = orth_rows(W_k^H B_c) -> all-gather (keep x n)

    Returns ``(U, S, Vh, keep)`` replicated on every rank, or None (on every rank) when the block must go through the
    gathered rank-0 split: compression rejected, a Cholesky breakdown, or a recent rejection of this shape."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mr, n = theta_rows.shape
    assert mr * world == m and n % world == 0
    key = (m, n, int(maxdim))
    if _backoff_dist.get(key, 0) > 0:
        _backoff_dist[key] -= 1
        stats["sketch_skipped"] += 1
        return None
    dt, dev = theta_rows.dtype, theta_rows.device
    rdt = theta_rows.real.dtype if theta_rows.is_complex() else dt
    k = min(int(maxdim), m, n)
    l = k + _oversample(k)
    nc = n // world
    c0 = rank * nc

    def count(name):
        if comm is not None:
            comm[name] = comm.get(name, 0) + 1

    def allsum(t):
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, group=group)
        count("all_reduce_split")

    def agree(flag: bool) -> bool:
        f = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
        return bool(f.item() > 0)

    def reject():
        stats["sketch_rejected"] += 1
        _backoff_dist[key] = SKETCH_BACKOFF
        return None

    Y = _mm(theta_rows, _omega(n, l, dt, dev))

    def regularise(Yc, tag, rows_total, r0):
        if SKETCH_REG > 0:
            ny = (torch.linalg.norm(Yc) ** 2).reshape(1)
            allsum(ny)
            reg = _fixed_normal(tag, rows_total, l, dt, dev, 0xB200)[r0:r0 + Yc.shape[0]]
            Yc.add_(reg, alpha=float(SKETCH_REG * torch.sqrt(ny)) / (rows_total * l) ** 0.5)
        return Yc

    def compress(Yc):
        Qc = _orthonormal_cols(regularise(Yc, "reg", m, rank * mr), m_total=m, allsum=allsum, agree=agree)
        if Qc is None:
            return None, None, None
        Bc_ = _mm(Qc.T, theta_rows, conj_a=True)               # partial l x n
        allsum(Bc_)
        w_ = torch.stack([torch.linalg.norm(theta_rows - _mm(Qc, Bc_)) ** 2, torch.linalg.norm(theta_rows) ** 2])
        allsum(w_)
        return Qc, Bc_, w_
    Q, B, w = compress(Y)
    if Q is None:
        return reject()
    stats["last_rho"] = float(w[0] / w[1])
    if SKETCH_TOL < stats["last_rho"] <= SKETCH_REFINE_BELOW:  # one decision for all ranks (all-reduced rho)
        # subspace-iteration refinement, see _sketch_split: V = orth(B^H) with the rows of B^H (columns of theta) sharded
        stats["sketch_refined"] = stats.get("sketch_refined", 0) + 1
        Bh = B[:, c0:c0 + nc]
        Bh = Bh.conj().T.resolve_conj().contiguous()                          # this rank's nc rows of B^H (n x l)
        Vc = _orthonormal_cols(regularise(Bh, "reg2", n, c0), m_total=n, allsum=allsum, agree=agree)
        if Vc is None:
            return reject()
        V1 = torch.empty(n, l, dtype=dt, device=dev)
        dist.all_gather_into_tensor(V1.view(-1), Vc.contiguous().view(-1), group=group)
        Q, B, w = compress(_mm(theta_rows, V1))
        if Q is None:
            return reject()
        stats["last_rho"] = float(w[0] / w[1])
    rho = w[0]
    if not stats["last_rho"] <= SKETCH_TOL:
        return reject()
    Bc = B[:, c0:c0 + nc]
    Gs = _mm(Bc, Bc.T, conj_b=True)
    allsum(Gs)
    meta = torch.zeros(1, dtype=torch.int64, device=dev)
    S = torch.empty(l + 1, dtype=rdt, device=dev)
    if rank == 0:
        Gs = 0.5 * (Gs + Gs.conj().T)
        lam, vecs = torch.linalg.eigh(Gs)
        lam = torch.flip(lam, dims=(0,)).clamp_min(0)
        S = torch.cat([torch.sqrt(lam), torch.sqrt(rho).reshape(1).to(rdt)]).contiguous()
        meta[0] = _keep_from_spectrum(S, k, cutoff)
    src = dist.get_global_rank(group, 0)
    dist.broadcast(meta, src=src, group=group)
    dist.broadcast(S, src=src, group=group)
    keep = int(meta[0])
    if rank == 0:
        top = torch.flip(vecs[:, vecs.shape[1] - keep:], dims=(1,)).contiguous()
    else:
        top = torch.empty(l, keep, dtype=dt, device=dev)
    dist.broadcast(torch.view_as_real(top) if top.is_complex() else top, src=src, group=group)
    count("broadcast_split")
    Ur = _mm(Q, top).contiguous()                              # mr x keep
    U = torch.empty(m, keep, dtype=dt, device=dev)
    dist.all_gather_into_tensor(U.view(-1), Ur.view(-1), group=group)
    Bk = _mm(top.T, Bc, conj_a=True)                           # keep x nc : this rank's columns of S_k Vh_k
    Vc = _orthonormal_rows(Bk, allsum=allsum, agree=agree)
    if Vc is None:
        return reject()
    buf = torch.empty(world, keep, nc, dtype=dt, device=dev)
    dist.all_gather_into_tensor(buf.view(-1), Vc.contiguous().view(-1), group=group)
    count("all_gather_split")
    Vh = buf.permute(1, 0, 2).reshape(keep, n)
    stats["sketch_accepted"] += 1
    return U, S, Vh, keep


def gram_split_sharded(theta_rows: torch.Tensor, m: int, maxdim: int, cutoff: float, group, comm: dict = None):
    """The full Gram/eigh split (``_gram_split``, right-hand Gram matrix) on a ROW-SHARDED theta -- the exact path for
    blocks whose truncation is real (the range compression rejects them) without gathering theta:

      G   = sum_r theta_r^H theta_r            local GEMM + all-reduce (n x n)
      eigh(G)                                  rank 0; spectrum and the leading eigenvectors V_k broadcast
      Vh  = V_k^H                              replicated, rows orthonormal
      (U S)_r = theta_r V_k                    local GEMM; columns orthogonal, norms S_j
      U_r = orth_cols((U S)_r)                 CholeskyQR2 on the normalised columns, k x k Gram matrices all-reduced
      U   = all-gather of the U_r

    Requires n <= m (the smaller Gram matrix is the right-hand one) and m divisible by the world size.  Returns
    ``(U, S, Vh, keep)`` replicated on every rank, or None (on every rank) if a Cholesky factorisation breaks down."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mr, n = theta_rows.shape
    assert mr * world == m and n <= m
    dt, dev = theta_rows.dtype, theta_rows.device
    rdt = theta_rows.real.dtype if theta_rows.is_complex() else dt

    def allsum(t):
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, group=group)
        if comm is not None:
            comm["all_reduce_split"] = comm.get("all_reduce_split", 0) + 1

    def agree(flag: bool) -> bool:
        f = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
        return bool(f.item() > 0)

    G = _mm(theta_rows.T, theta_rows, conj_a=True)             # n x n partial
    allsum(G)
    src = dist.get_global_rank(group, 0)
    meta = torch.zeros(1, dtype=torch.int64, device=dev)
    S = torch.empty(n, dtype=rdt, device=dev)
    if rank == 0:
        G = 0.5 * (G + G.conj().T)
        lam, vecs = torch.linalg.eigh(G)
        S = torch.sqrt(torch.flip(lam, dims=(0,)).clamp_min(0)).contiguous()
        meta[0] = _keep_from_spectrum(S, maxdim, cutoff)
    del G
    dist.broadcast(meta, src=src, group=group)
    dist.broadcast(S, src=src, group=group)
    keep = int(meta[0])
    if rank == 0:
        top = torch.flip(vecs[:, vecs.shape[1] - keep:], dims=(1,)).contiguous()        # n x keep
        del vecs
    else:
        top = torch.empty(n, keep, dtype=dt, device=dev)
    dist.broadcast(torch.view_as_real(top) if top.is_complex() else top, src=src, group=group)
    if comm is not None:
        comm["broadcast_split"] = comm.get("broadcast_split", 0) + 1
    US = _mm(theta_rows, top)                                  # mr x keep
    UrH = _orthonormal_rows(US.conj().T.resolve_conj().contiguous(), allsum=allsum, agree=agree)
    if UrH is None:
        return None
    Ur = UrH.conj().T.resolve_conj().contiguous()
    U = torch.empty(m, keep, dtype=dt, device=dev)
    dist.all_gather_into_tensor(U.view(-1), Ur.view(-1), group=group)
    if comm is not None:
        comm["all_gather_split"] = comm.get("all_gather_split", 0) + 1
    stats["gram_sharded"] = stats.get("gram_sharded", 0) + 1
    return U, S, top.conj().T.resolve_conj().contiguous(), keep


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


def _oversample(k: int) -> int:
    return max(16, min(SKETCH_OVERSAMPLE, k // 8))


TIMING = False        # True: synchronise around every large split and accumulate wall ms per path in `stats`


def _large_split(theta: torch.Tensor, maxdim: int, cutoff: float, sketch: bool = None):
    if not (TIMING and theta.is_cuda):
        return _large_split_impl(theta, maxdim, cutoff, sketch)
    import time
    snap = (stats["sketch_accepted"], stats["sketch_rejected"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = _large_split_impl(theta, maxdim, cutoff, sketch)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    path = ("ms_sketch_accepted" if stats["sketch_accepted"] != snap[0] else
            "ms_sketch_rejected_then_gram" if stats["sketch_rejected"] != snap[1] else "ms_gram")
    stats[path] = stats.get(path, 0.0) + ms
    return out


def _large_split_impl(theta: torch.Tensor, maxdim: int, cutoff: float, sketch: bool = None):
    small = min(theta.shape)
    if (SKETCH if sketch is None else sketch) and int(maxdim) + _oversample(int(maxdim)) <= SKETCH_MAX_FRACTION * small:
        key = (tuple(theta.shape), int(maxdim))
        if _backoff.get(key, 0) > 0:
            _backoff[key] -= 1
            stats["sketch_skipped"] += 1
        else:
            out = _sketch_split(theta, maxdim, cutoff)
            if out is not None:
                stats["sketch_accepted"] += 1
                return out
            stats["sketch_rejected"] += 1
            _backoff[key] = SKETCH_BACKOFF
    stats["gram"] += 1
    return _gram_split(theta, maxdim, cutoff)


def two_site_split(theta: torch.Tensor, maxdim: int, cutoff: float, *, fast_min: int = None, sketch: bool = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """Returns ``(U[:, :keep], S_all, Vh[:keep, :], keep)`` for the 2-D block ``theta`` (chi_l*d, d*chi_r).
    ``S_all`` is the full descending spectrum (the caller records the discarded weight from ``S_all[keep:]``)."""
    fm = FAST_MIN if fast_min is None else fast_min
    if min(theta.shape) >= fm:
        if theta.is_cuda:
            from ._lib import nvtx_range
            with nvtx_range("qsb.two_site_split"):
                return _large_split(theta, maxdim, cutoff, sketch)
        return _large_split(theta, maxdim, cutoff, sketch)
    stats["svd"] += 1
    U, S, Vh = torch.linalg.svd(theta, full_matrices=False)         # dmrg.py:685
    keep = _keep_from_spectrum(S, maxdim, cutoff)
    return U[:, :keep].resolve_conj(), S, Vh[:keep, :].resolve_conj(), keep

"""The two-site split (svd.py) is plain torch and device-agnostic: its Gram/eigh path is checked on the CPU against
torch.linalg.svd and the reference's truncation rule (dmrg.py:687-695)."""
import pytest
import torch

from quantum_simulation_b200.svd import two_site_split, _keep_from_spectrum


def _matrix(m, n, dtype, decay, seed):
    g = torch.Generator().manual_seed(seed)
    r = min(m, n)
    U0, _ = torch.linalg.qr(torch.randn(m, r, dtype=dtype, generator=g))
    V0, _ = torch.linalg.qr(torch.randn(n, r, dtype=dtype, generator=g))
    s = torch.exp(-torch.arange(r, dtype=torch.float64) * decay)
    return (U0 * s.to(dtype)) @ V0.conj().T, s


def _reference_keep(S, maxdim, cutoff):
    keep_max = min(S.numel(), maxdim)
    keep_cut = S.numel()
    if cutoff > 0.0:
        S2 = S ** 2
        cs = torch.cumsum(S2 / torch.sum(S2), dim=0)
        keep_cut = int(torch.searchsorted(cs, 1.0 - cutoff, right=True)) + 1
    return max(1, min(keep_max, keep_cut))


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(96, 96), (64, 160), (160, 64), (1, 8), (8, 1), (2, 2)])
@pytest.mark.parametrize("maxdim,cutoff", [(10 ** 6, 0.0), (24, 0.0), (10 ** 6, 1e-10), (5, 1e-6)])
def test_gram_split_matches_svd(dtype, shape, maxdim, cutoff, monkeypatch):
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "SKETCH", False)                 # the full Gram/eigh path (the sketch has its own tests)
    m, n = shape
    theta, _ = _matrix(m, n, dtype, 0.2, seed=m * 1000 + n)
    U, S, Vh, keep = two_site_split(theta, maxdim, cutoff, fast_min=1)          # force the Gram/eigh path
    Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
    assert keep == _reference_keep(Se, maxdim, cutoff) == _keep_from_spectrum(Se, maxdim, cutoff)
    assert U.shape == (m, keep) and Vh.shape == (keep, n) and S.shape == Se.shape
    eye = torch.eye(keep, dtype=dtype)
    assert float(torch.linalg.norm(U.conj().T @ U - eye)) < 1e-12
    assert float(torch.linalg.norm(Vh @ Vh.conj().T - eye)) < 1e-12
    # S_i = sqrt(eigenvalue): absolute accuracy eps * S_max^2 / S_i (documented in svd.py), i.e. exact to ~1e-15 at
    # the top of the spectrum, relative 1e-4 for S_i / S_max = 1e-6, noise below sqrt(eps) * S_max
    top = Se > 1e-7 * Se[0]
    assert bool(((S - Se).abs()[top] <= 2e-14 * Se[0] ** 2 / Se[top]).all())
    assert float((S[:4] - Se[:4]).abs().max()) <= 1e-13 * float(Se[0])
    rec = (U * S[:keep].to(dtype)) @ Vh
    best = (Ue[:, :keep] * Se[:keep].to(dtype)) @ Vhe[:keep]
    assert float(torch.linalg.norm(rec - best)) < 5e-8 * float(Se[0])
    # the small-block path is torch.linalg.svd itself
    U2, S2, Vh2, keep2 = two_site_split(theta, maxdim, cutoff, fast_min=10 ** 9)
    assert keep2 == keep and torch.allclose(S2, Se)


def test_gram_split_rank_deficient_and_degenerate(monkeypatch):
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "SKETCH", False)
    _rank_deficient_and_degenerate(1e-12)


def test_sketched_split_rank_deficient_and_degenerate(monkeypatch):
    """Same blocks through the range compression: reconstruction to the regularisation amplitude (1e-9 relative)."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "_backoff", {})
    before = qsvd.stats["sketch_accepted"]
    _rank_deficient_and_degenerate(5e-9)
    assert qsvd.stats["sketch_accepted"] >= before + 2


def _rank_deficient_and_degenerate(rec_tol):
    """Exactly degenerate and exactly zero singular values (product states, symmetry multiplets): still isometric
    factors and the right weights; the arbitrary rotation inside a degenerate multiplet does not matter."""
    g = torch.Generator().manual_seed(3)
    Q1, _ = torch.linalg.qr(torch.randn(40, 40, dtype=torch.float64, generator=g))
    Q2, _ = torch.linalg.qr(torch.randn(48, 40, dtype=torch.float64, generator=g))
    s = torch.tensor([1.0] * 3 + [0.5] * 4 + [0.0] * 33, dtype=torch.float64)
    theta = (Q1 * s) @ Q2.T
    U, S, Vh, keep = two_site_split(theta, 7, 0.0, fast_min=1)
    assert keep == 7
    assert torch.allclose(S[:7], s[:7], atol=1e-12) and float(S[7:].max()) < 1e-7
    assert float(torch.linalg.norm(U.T @ U - torch.eye(7, dtype=torch.float64))) < 1e-12
    assert float(torch.linalg.norm(Vh @ Vh.T - torch.eye(7, dtype=torch.float64))) < 1e-12
    assert float(torch.linalg.norm((U * S[:7]) @ Vh - theta)) < rec_tol
    U, S, Vh, keep = two_site_split(theta, 100, 1e-12, fast_min=1)              # cutoff drops the numerical zeros
    assert keep == 7
    rank1 = torch.outer(Q1[:, 0], Q2[:, 0])
    U, S, Vh, keep = two_site_split(rank1, 4, 0.0, fast_min=1)
    assert keep == 4 and abs(float(S[0]) - 1.0) < 1e-12
    assert float(torch.linalg.norm(U.T @ U - torch.eye(4, dtype=torch.float64))) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape,k,decay", [((256, 256), 64, 0.3), ((256, 320), 64, 0.25), ((320, 256), 64, 0.25),
                                           ((192, 192), 40, 0.5)])
def test_sketched_split_accepted_when_the_tail_is_negligible(dtype, shape, k, decay, monkeypatch):
    """Range compression in front of the Gram/eigh path (svd._sketch_split): taken for truncating blocks, accepted only
    when theta has <= 1e-13 of its weight outside the compressed range, and then as accurate as the full path: squared
    Schmidt values within rho, isometric factors, the optimally truncated block reproduced to sqrt(rho)."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "_backoff", {})
    m, n = shape
    theta, _ = _matrix(m, n, dtype, decay, seed=m + n + k)
    theta = theta / torch.linalg.norm(theta)
    before = dict(qsvd.stats)
    U, S, Vh, keep = two_site_split(theta, k, 0.0, fast_min=1)
    assert qsvd.stats["sketch_accepted"] == before["sketch_accepted"] + 1 and qsvd.stats["last_rho"] <= 1e-13
    Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
    assert keep == k and U.shape == (m, k) and Vh.shape == (k, n)
    l = k + qsvd._oversample(k)
    assert S.numel() == l + 1                                  # leading l values + the residual pseudo-value
    assert float((S[:l] ** 2 - Se[:l] ** 2).abs().max()) <= 2e-13
    assert abs(float((S[keep:] ** 2).sum()) - float((Se[keep:] ** 2).sum())) <= 2e-13      # discarded weight
    eye = torch.eye(k, dtype=dtype)
    assert float(torch.linalg.norm(U.conj().T @ U - eye)) < 1e-12
    assert float(torch.linalg.norm(Vh @ Vh.conj().T - eye)) < 1e-12
    rec = (U * S[:keep].to(dtype)) @ Vh
    best = (Ue[:, :keep] * Se[:keep].to(dtype)) @ Vhe[:keep]
    assert float(torch.linalg.norm(rec - best)) < 1e-6
    # a cutoff that bites inside the kept range gives the reference's keep
    U2, S2, Vh2, keep2 = two_site_split(theta, k, 1e-9, fast_min=1)
    assert keep2 == _reference_keep(Se, k, 1e-9)


def test_sketched_split_rejected_when_the_tail_matters(monkeypatch):
    """Slowly decaying spectrum: the compressed range misses weight >> 1e-13, the sketch is rejected, the full Gram path
    answers (identical to calling it directly), and the next blocks of that shape skip the attempt."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "_backoff", {})
    theta, _ = _matrix(256, 256, torch.float64, 0.02, seed=9)
    before = dict(qsvd.stats)
    U, S, Vh, keep = two_site_split(theta, 64, 0.0, fast_min=1)
    assert qsvd.stats["sketch_rejected"] == before["sketch_rejected"] + 1 and qsvd.stats["last_rho"] > 1e-13
    Ug, Sg, Vhg, kg = qsvd._gram_split(theta, 64, 0.0)
    assert keep == kg and torch.equal(S, Sg) and torch.equal(U, Ug) and torch.equal(Vh, Vhg)
    two_site_split(theta, 64, 0.0, fast_min=1)
    assert qsvd.stats["sketch_skipped"] == before["sketch_skipped"] + 1
    assert qsvd.stats["sketch_rejected"] == before["sketch_rejected"] + 1


def test_sketched_split_rank_deficient_block():
    """Exact low rank (first sweeps from a product-like state): the shifted CholeskyQR3 breaks down or flags itself,
    Householder QR takes over, the result is still the exact split."""
    from quantum_simulation_b200 import svd as qsvd
    g = torch.Generator().manual_seed(5)
    A = torch.randn(256, 12, dtype=torch.float64, generator=g)
    Bm = torch.randn(12, 256, dtype=torch.float64, generator=g)
    theta = A @ Bm
    theta = theta / torch.linalg.norm(theta)
    U, S, Vh, keep = two_site_split(theta, 64, 0.0, fast_min=1)
    Se = torch.linalg.svdvals(theta)
    assert keep == 64 and float((S[:12] - Se[:12]).abs().max()) < 1e-13
    assert float(torch.linalg.norm((U * S[:keep]) @ Vh - theta)) < 1e-7
    eye = torch.eye(64, dtype=torch.float64)
    assert float(torch.linalg.norm(U.T @ U - eye)) < 1e-12 and float(torch.linalg.norm(Vh @ Vh.T - eye)) < 1e-12


@pytest.mark.parametrize("decay", [0.03, 0.05, 0.08])
def test_cutoff_driven_keep_matches_svd_at_the_default_cutoff(decay, monkeypatch):
    """the Gram path's eigenvalues carry absolute noise ~eps * lambda_max, the same order as the reference's
    default cutoff (Sweeps.cutoff = 1e-12) once summed over a long tail.  On graded (DMRG-like) spectra the kept bond
    dimension and the discarded weight still agree with the SVD path: the noise is per-eigenvalue 1e-16 with random
    sign, so the tail sum moves by ~sqrt(#tail) * 1e-16 << cutoff.  Both the full and the range-compressed path."""
    from quantum_simulation_b200 import svd as qsvd
    for seed in range(4):
        theta, _ = _matrix(384, 384, torch.float64, decay, seed=seed)
        theta = theta / torch.linalg.norm(theta)
        Se = torch.linalg.svdvals(theta)
        for cutoff in (1e-11, 1e-12, 1e-13):
            kr = _reference_keep(Se, 10 ** 6, cutoff)
            for sketch, maxdim in ((False, 10 ** 6), (True, 150)):
                monkeypatch.setattr(qsvd, "_backoff", {})
                U, S, Vh, keep = two_site_split(theta, maxdim, cutoff, fast_min=1, sketch=sketch)
                assert keep == min(kr, maxdim)
                assert abs(float((S[keep:] ** 2).sum()) - float((Se[keep:] ** 2).sum())) <= 5e-14


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
def test_sketched_split_refines_a_marginal_miss(dtype, monkeypatch):
    """rho of the plain compression a few times above the bound: one subspace-iteration step with re-orthonormalisation
    brings it to ~1.6 x the weight beyond the compressed dimension, and the block is accepted."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "_backoff", {})
    monkeypatch.setattr(qsvd, "SKETCH_REFINE_SINGLE", True)
    theta, _ = _matrix(256, 256, dtype, 0.2, seed=1)
    theta = theta / torch.linalg.norm(theta)
    before = dict(qsvd.stats)
    U, S, Vh, keep = two_site_split(theta, 64, 0.0, fast_min=1)
    assert qsvd.stats.get("sketch_refined", 0) == before.get("sketch_refined", 0) + 1
    assert qsvd.stats["sketch_accepted"] == before["sketch_accepted"] + 1 and qsvd.stats["last_rho"] <= 1e-13
    Se = torch.linalg.svdvals(theta)
    assert float((S[:64] ** 2 - Se[:64] ** 2).abs().max()) <= 2e-13
    eye = torch.eye(64, dtype=dtype)
    assert float(torch.linalg.norm(U.conj().T @ U - eye)) < 1e-12 and float(torch.linalg.norm(Vh @ Vh.conj().T - eye)) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape,cutoff", [((512, 512), 0.0), ((512, 640), 0.0), ((640, 512), 0.0), ((512, 512), 1e-12)])
def test_rank_adaptive_compression_pads_an_over_resolved_bond(dtype, shape, cutoff, monkeypatch):
    """An over-resolved block (numerical rank ~16, bond dimension 192): after the first split of the shape has recorded
    the rank, the next ones are compressed to 64 columns only, the 64 x 64 eigenproblem replaces the 216 x 216 one and
    the kept basis is padded with weightless orthonormal directions up to the reference's bond dimension -- same keep,
    isometric factors, the optimally truncated block to the regularisation amplitude, exact discarded weight."""
    from quantum_simulation_b200 import svd as qsvd
    qsvd.clear_caches()
    m, n = shape
    theta, _ = _matrix(m, n, dtype, 1.2, seed=3)
    theta = theta / torch.linalg.norm(theta)
    Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
    before = qsvd.stats.get("sketch_narrow", 0)
    for rep in range(3):
        U, S, Vh, keep = two_site_split(theta, 192, cutoff, fast_min=1)
    assert qsvd.stats.get("sketch_narrow", 0) == before + 2 and qsvd.stats["last_rho"] <= 1e-15
    kr = _reference_keep(Se, 192, cutoff)
    assert keep == kr and U.shape == (m, keep) and Vh.shape == (keep, n)
    eye = torch.eye(keep, dtype=dtype)
    assert float(torch.linalg.norm(U.conj().T @ U - eye)) < 1e-12
    assert float(torch.linalg.norm(Vh @ Vh.conj().T - eye)) < 1e-12
    best = (Ue[:, :kr] * Se[:kr].to(dtype)) @ Vhe[:kr]
    assert float(torch.linalg.norm((U * S[:keep].to(dtype)) @ Vh - best)) < 1e-7
    assert abs(float((S[keep:] ** 2).sum()) - float((Se[kr:] ** 2).sum())) <= 1e-15
    assert float((S[:8] - Se[:8]).abs().max()) < 1e-13
    # the block outgrows the hint: retried at full width, still right
    wide, _ = _matrix(m, n, dtype, 0.12, seed=4)
    wide = wide / torch.linalg.norm(wide)
    r0 = qsvd.stats.get("sketch_narrow_retry", 0)
    U, S, Vh, keep = two_site_split(wide, 192, cutoff, fast_min=1)
    assert qsvd.stats.get("sketch_narrow_retry", 0) == r0 + 1
    Sw = torch.linalg.svdvals(wide)
    assert keep == _reference_keep(Sw, 192, cutoff) and float((S[:8] - Sw[:8]).abs().max()) < 1e-13

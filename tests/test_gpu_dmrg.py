"""End-to-end GPU parity: DMRG(...)(sweeps) through our drop-in classes against the per-update energies,
bond dimensions and discarded weights recorded from the unmodified reference (tests/golden/dmrg.npz).
Tolerance (BASELINE.json): energy within 1e-10 relative, checked at EVERY local update, not just per sweep."""
import io
import os
import contextlib

import numpy as np
import pytest
import torch

from conftest import load_golden, golden_mpo
import quantum_simulation_b200 as qsb

pytestmark = pytest.mark.gpu
DEV = "cuda"


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_case(name):
    z = load_golden("dmrg")
    L, d, chi0, seed, ns = (int(x) for x in z[f"{name}/cfg"])
    dt = torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64
    mps = quiet(qsb.MPS, L, d, chi0, "cpu", dt, seed).to(DEV)       # CPU-seeded like the reference, then moved
    mpo = golden_mpo(str(z[f"{name}/mpo"]), dt, DEV)
    sw = qsb.Sweeps(ns)
    sw.set_schedule(maxdim=[int(x) for x in z[f"{name}/maxdim"]], krylov_dim=[int(x) for x in z[f"{name}/krylov_dim"]],
                    cutoff=[float(x) for x in z[f"{name}/cutoff"]], reortho=[bool(x) for x in z[f"{name}/reortho"]],
                    maxreortho=[int(x) for x in z[f"{name}/maxreortho"]], noise=[0.0] * ns)
    dm = quiet(qsb.DMRG, mps, mpo, device=DEV, dtype=dt, contract_backend="ncon")
    E, out_mps = quiet(dm, sw, dispon=0, maxit=2)
    return z, np.array(E), dm, out_mps


@pytest.mark.parametrize("name", [str(s) for s in load_golden("dmrg")["names"]])
def test_dmrg_matches_reference_per_update(name):
    z, E, dm, mps = run_case(name)
    Eg = z[f"{name}/energies"]
    assert len(E) == len(Eg)
    assert dm.history["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    rel = np.max(np.abs(E - Eg) / np.abs(Eg))
    assert rel < 1e-10, f"max relative energy deviation {rel:.2e}"
    dw = np.array(dm.history["discarded_weights"])
    assert np.allclose(dw, z[f"{name}/discarded"], rtol=1e-6, atol=1e-12)
    assert set(dm.history) == {"energies", "discarded_weights", "bond_dims", "sweeps", "directions", "sites"}
    assert mps.ortho_center == 0


def test_notebook_energy_printout_format(capsys):
    z = load_golden("dmrg")
    mps = qsb.MPS(6, 2, 4, "cpu", torch.complex128, 2024).to(DEV)
    mpo = golden_mpo("tfim_c128_L6_g0.7", torch.complex128, DEV)
    sw = qsb.Sweeps(1)
    sw.set_schedule(maxdim=[8])
    dm = qsb.DMRG(mps, mpo, device=DEV, dtype=torch.complex128)
    dm(sw, dispon=1)
    out = capsys.readouterr().out
    assert "Initializing MPS as a right-canonical random state..." in out
    assert "[DMRG] Contraction backend resolved to" in out
    assert "Sweep: 1 of 1, Energy: -2.3275" in out and "Bond dim:" in out


# ----------------------------------------------------------------------------- two-site split ("next" row f.1)
@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(512, 512), (384, 640), (640, 384)])
def test_two_site_split_gram_path(dtype, shape):
    """The Gram/eigh split used for large blocks: isometric factors, the reference's truncation rule, the same
    kept subspace and spectrum as torch.linalg.svd."""
    from quantum_simulation_b200.svd import two_site_split
    m, n = shape
    g = torch.Generator().manual_seed(3)
    r = min(m, n)
    U0, _ = torch.linalg.qr(torch.randn(m, r, dtype=dtype, generator=g))
    V0, _ = torch.linalg.qr(torch.randn(n, r, dtype=dtype, generator=g))
    s = torch.exp(-torch.arange(r, dtype=torch.float64) * 0.05)
    theta = ((U0 * s.to(dtype)) @ V0.conj().T).to(DEV)
    for maxdim, cutoff in ((200, 0.0), (10 ** 6, 1e-10)):
        U, S, Vh, keep = two_site_split(theta, maxdim, cutoff)
        Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
        S2 = Se ** 2
        keep_ref = min(r, maxdim)
        if cutoff > 0:
            cs = torch.cumsum(S2 / S2.sum(), 0)
            keep_ref = min(keep_ref, int(torch.searchsorted(cs, torch.tensor(1.0 - cutoff, device=DEV, dtype=cs.dtype), right=True)) + 1)
        assert keep == keep_ref
        eye = torch.eye(keep, dtype=dtype, device=DEV)
        assert float(torch.linalg.norm(U.conj().T @ U - eye)) < 1e-12
        assert float(torch.linalg.norm(Vh @ Vh.conj().T - eye)) < 1e-12
        assert float(((S[:keep] - Se[:keep]).abs() / Se[:keep]).max()) < 1e-6
        rec = (U * S[:keep].to(dtype)) @ Vh
        best = (Ue[:, :keep] * Se[:keep].to(dtype)) @ Vhe[:keep]
        assert float(torch.linalg.norm(rec - best)) < 1e-7 * float(Se[0])


@pytest.mark.parametrize("name", ["cal_heis_L20", "c1_xxz_L20_chi64_c128", "heis_f64_L20_pro", "nb02_tfim_L6"])
def test_dmrg_with_gram_split_everywhere(name, monkeypatch):
    """Force the Gram/eigh split for EVERY block size: per-update energies and bond dimensions still match the
    reference (energy within 1e-10 relative)."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "FAST_MIN", 2)
    z, E, dm, mps = run_case(name)
    Eg = z[f"{name}/energies"]
    assert dm.history["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    assert np.max(np.abs(E - Eg) / np.abs(Eg)) < 1e-10


def test_baseline_config2_first_sweep_against_reference():
    """BASELINE config 2 at full size: TFIM L=100, g=0.5, chi=512, complex128, sweep 1 (198 local updates) against
    the per-update energies of the unmodified reference (tests/golden/dmrg_c2_chi512.npz, 315 s on 8 CPU cores).

    From a RANDOM chi=512 start the first sweep is not reproducible to 1e-10 by anyone: the reference and its own
    op-for-op restatement on the same CPU differ by 1.7e-4 at the first (tiny, near-breakdown) update and by 7e-8 at
    the end of the sweep (stored as oracle_energies).  The bar here is therefore that envelope, not 1e-10; the 1e-10
    per-update bar is enforced on every case where the reference reproduces itself (test_dmrg_matches_reference_per_update)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dmrg_c2_chi512.npz"))
    Eg, Eo = z["energies"], z["oracle_energies"]
    mps = quiet(qsb.MPS, 100, 2, 512, "cpu", torch.complex128, 1)
    sums = np.array([t.detach().abs().sum().item() for t in mps.B])
    assert np.allclose(sums, z["init_abs_sums"], rtol=1e-13, atol=0)
    mps = mps.to(DEV)
    sw = qsb.Sweeps(1)
    sw.set_schedule(maxdim=[512], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
    dm = quiet(qsb.DMRG, mps, golden_mpo("tfim_c128_L100_g0.5", torch.complex128, DEV), device=DEV,
               dtype=torch.complex128)
    E, _ = quiet(dm, sw, dispon=0, maxit=2)
    E = np.array(E)
    assert dm.history["bond_dims"] == [int(x) for x in z["bond_dims"]]
    rel = np.abs(E - Eg) / np.abs(Eg)
    ref_spread = np.abs(Eo - Eg) / np.abs(Eg)
    assert rel[:20].max() < 2e-3 and ref_spread[:20].max() < 2e-3
    # the reference-vs-restatement spread over updates >= 20 is 4e-7; allow the same order of magnitude
    assert rel[20:].max() < 5e-6, f"deviation {rel[20:].max():.2e} (reference-vs-restatement spread {ref_spread[20:].max():.2e})"
    assert abs(E[-1] - Eg[-1]) / abs(Eg[-1]) < 1e-6


# ----------------------------------------------------------------------------- environment-based observables (f.4)
def _dense_mpo(mpo):
    """Dense matrix of an open-boundary MPO (small systems): the checker for the environment contractions."""
    acc = mpo[0][0]                                     # (Wr, d, d) after fixing the left boundary index
    for M in mpo[1:]:
        acc = torch.einsum("wab,wvcd->vacbd", acc, M)
        v, a, c, b, d = acc.shape
        acc = acc.reshape(v, a * c, b * d)
    return acc[0]


@pytest.mark.parametrize("name", ["nb02_tfim_L6", "nb03_heis_L4"])
def test_energy_variance_by_environments(name):
    """<H> and Var(H) from left-environment contractions against the dense d^N computation the reference uses
    (dmrg_observables.py:539-582), on converged notebook runs (dmrg_04 stores Var(H) = 1.35e-12 for TFIM L=6)."""
    z, E, dm, mps = run_case(name)
    dt = mps.B[0].dtype
    mpo = golden_mpo(str(z[f"{name}/mpo"]), dt, DEV)
    state = mps.B[0].detach()
    for t in list(mps.B)[1:]:
        state = torch.tensordot(state, t.detach(), dims=([-1], [0]))
    psi = state.reshape(-1)
    H = _dense_mpo(mpo)
    Hpsi = H @ psi
    nrm = torch.vdot(psi, psi).real
    e_dense = (torch.vdot(psi, Hpsi).real / nrm).item()
    var_dense = (torch.vdot(Hpsi, Hpsi).real / nrm).item() - e_dense ** 2
    e_env = float(qsb.energy_expectation(mps, mpo))
    var_env = float(qsb.energy_variance(mps, mpo))
    assert abs(e_env - e_dense) <= 1e-12 * max(1.0, abs(e_dense))
    assert abs(e_env - E[-1]) <= 1e-9                      # the last Ritz value is the energy of the final state
    assert abs(var_env - max(var_dense, 0.0)) <= 1e-11
    assert 0.0 <= var_env < 1e-9
    with pytest.raises(ValueError):
        qsb.energy_variance(mps, mpo, form="C")
    with pytest.raises(ValueError):
        qsb.energy_variance(mps, mpo[:-1])


def test_energy_variance_unconverged_state_matches_dense():
    """A random (unconverged) MPS has a large variance: checks <H^2> through the W^2-bond squared MPO."""
    dt = torch.complex128
    mps = quiet(qsb.MPS, 8, 2, 12, "cpu", dt, 21).to(DEV)
    mpo = golden_mpo("heis_c128_L20", dt, DEV)
    mpo8 = [mpo[0], mpo[1]] + [mpo[5]] * 4 + [mpo[-2], mpo[-1]]
    state = mps.B[0].detach()
    for t in list(mps.B)[1:]:
        state = torch.tensordot(state, t.detach(), dims=([-1], [0]))
    psi = state.reshape(-1)
    H = _dense_mpo(mpo8)
    Hpsi = H @ psi
    nrm = torch.vdot(psi, psi).real
    e_dense = (torch.vdot(psi, Hpsi).real / nrm).item()
    var_dense = (torch.vdot(Hpsi, Hpsi).real / nrm).item() - e_dense ** 2
    assert abs(float(qsb.energy_expectation(mps, mpo8)) - e_dense) <= 1e-12 * max(1.0, abs(e_dense))
    assert abs(float(qsb.energy_variance(mps, mpo8)) - var_dense) <= 1e-11 * max(1.0, var_dense)
    assert var_dense > 1e-3


def test_mps_built_on_the_gpu_is_right_canonical_and_drives_dmrg():
    """'Next' row f.3: the MPS mirror canonicalises directly on the GPU (torch QR on device) -- no CPU round trip at
    large chi -- and the result is a valid start state for DMRG (energy decreases monotonically sweep to sweep)."""
    dt = torch.float64
    mps = quiet(qsb.MPS, 20, 2, 32, DEV, dt, 4)
    assert mps.B[0].device.type == "cuda" and mps.ortho_center == 0
    for p in range(1, 20):
        b = mps.B[p].reshape(mps.B[p].shape[0], -1)
        assert float(torch.linalg.norm(b @ b.T - torch.eye(b.shape[0], dtype=dt, device=DEV))) < 1e-12
    assert abs(float(torch.linalg.norm(mps.B[0])) - 1.0) < 1e-13
    mpo = golden_mpo("heis_f64_L20", dt, DEV)
    sw = qsb.Sweeps(3)
    sw.set_schedule(maxdim=[32] * 3, cutoff=[0.0] * 3)
    dm = quiet(qsb.DMRG, mps, mpo, device=DEV, dtype=dt)
    E, _ = quiet(dm, sw, dispon=0)
    ends = [E[(k + 1) * 38 - 1] for k in range(3)]
    assert ends[0] >= ends[1] - 1e-12 >= ends[2] - 2e-12
    assert abs(ends[2] - (-8.682473319650)) < 1e-6          # SURVEY section 4 calibration value (chi = 32)
    assert abs(float(qsb.energy_expectation(dm.mps, mpo)) - ends[2]) < 1e-9


# ----------------------------------------------------------------------------- converged sweeps at BASELINE-size chains
# tests/golden/make_golden_sweeps.py: multi-sweep runs of the unmodified reference.  From a random start the first
# sweeps are chaotic even for the reference against itself (its 'ncon' and 'einsum' branches are 5e-5 apart after
# sweep 1, 2e-9 after sweep 2, <= 3e-14 from sweep 3 on -- stored in the fixtures), so the 1e-10 bar of BASELINE.json
# ("ground-state energy per sweep within 1e-10 relative") is asserted on every sweep from `from_sweep` on.
_SWEEP_CASES = {
    # fixture: (mpo key, first sweep (1-based) whose end-of-sweep energy is asserted at 1e-10)
    "c3s128": ("heis_f64_L100", 3),      # Heisenberg L=100 (config 3's chain), chi=128, f64, 6 sweeps
    "c3s": ("heis_f64_L100", 3),         # Heisenberg L=100, chi=256, f64, 4 sweeps
    "c2": ("tfim_c128_L100_g0.5", 7),    # BASELINE config 2 in full: TFIM L=100 g=0.5, chi=512, complex128, 10 sweeps
                                         # (the reference itself still moves by 5e-11 between sweeps 6 and 7)
}


def _run_sweeps_case(name):
    path = os.path.join(os.path.dirname(__file__), "golden", f"dmrg_sweeps_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"fixture {os.path.basename(path)} not generated")
    z = np.load(path)
    key, from_sweep = _SWEEP_CASES[name]
    L, d, chi, seed, ns = (int(x) for x in z["cfg"])
    dt = torch.complex128 if str(z["dtype"]) == "c128" else torch.float64
    mps = quiet(qsb.MPS, L, d, chi, "cpu", dt, seed)
    sums = np.array([t.detach().abs().sum().item() for t in mps.B])
    assert np.allclose(sums, z["init_abs_sums"], rtol=1e-13, atol=0)          # the reference's start state
    mps = mps.to(DEV)
    sw = qsb.Sweeps(ns)
    sw.set_schedule(maxdim=[chi] * ns, cutoff=[0.0] * ns, krylov_dim=[4] * ns, noise=[0.0] * ns)
    dm = quiet(qsb.DMRG, mps, golden_mpo(key, dt, DEV), device=DEV, dtype=dt)
    E, _ = quiet(dm, sw, dispon=0, maxit=2)
    E = np.array(E)
    per = len(E) // ns
    assert len(E) == len(z["energies"])
    assert dm.history["bond_dims"] == [int(x) for x in z["bond_dims"]]
    ends, ref_ends = E[per - 1::per], z["sweep_end"]
    rel = np.abs(ends - ref_ends) / np.abs(ref_ends)
    assert rel[from_sweep - 1:].max() < 1e-10, f"end-of-sweep relative deviation per sweep: {rel}"
    # every LOCAL update as well, one sweep later (during sweep `from_sweep` the runs are still converging onto each
    # other: the reference's own two branches are 1e-9 apart at its start and 1e-14 at its end)
    upd = np.abs(E - z["energies"]) / np.abs(z["energies"])
    if ns > from_sweep:
        assert upd[from_sweep * per:].max() < 1e-10, f"per-update deviation {upd[from_sweep * per:].max():.2e}"
    if "einsum_sweep_end" in z.files:        # the reference's own branch-to-branch spread: we must not be worse than 10x
        spread = np.abs(z["einsum_sweep_end"] - ref_ends) / np.abs(ref_ends)
        assert spread[from_sweep - 1:].max() < 1e-10
    return E, dm


@pytest.mark.parametrize("split", ["gram", "svd"])
@pytest.mark.parametrize("name", ["c3s128", "c3s"])
def test_converged_sweep_energies_config3_chain(name, split, monkeypatch):
    """Heisenberg L=100 (BASELINE config 3's chain) at a bond dimension the CPU reference can afford: end-of-sweep
    energies within 1e-10 relative of the unmodified reference from sweep 3 on, through both two-site-split paths."""
    from quantum_simulation_b200 import svd as qsvd
    monkeypatch.setattr(qsvd, "FAST_MIN", 2 if split == "gram" else 10 ** 9)
    _run_sweeps_case(name)


def test_converged_sweep_energies_config2_full_size():
    """BASELINE config 2 exactly as worded (TFIM L=100, chi=512, 10 sweeps, complex128): sweeps 7..10 within 1e-10
    relative of the unmodified reference's CPU run (tests/golden/dmrg_sweeps_c2.npz)."""
    E, dm = _run_sweeps_case("c2")
    mpo = golden_mpo("tfim_c128_L100_g0.5", torch.complex128, DEV)
    assert abs(float(qsb.energy_expectation(dm.mps, mpo)) - E[-1]) < 1e-9
    assert float(qsb.energy_variance(dm.mps, mpo)) < 1e-8


# ----------------------------------------------------------------------------- model families of BASELINE configs 4, 5
def _c45_mpo(z, name, dt):
    return [torch.from_numpy(z[f"{name}/mpo/t{int(u)}"]).to(dt).contiguous().to(DEV) for u in z[f"{name}/mpo/index"]]


@pytest.mark.parametrize("name", ["c4s_j1j2_4x4_f64", "c5s_hubbard_L10_c128"])
@pytest.mark.parametrize("split", ["default", "gram"])
def test_dmrg_config4_and_5_model_families(name, split, monkeypatch):
    """Full DMRG on the MODEL FAMILIES of BASELINE configs 4 and 5 at a size the CPU reference can afford (J1-J2 cylinder
    4 x 4: FSM MPO bonds up to 26; Fermi-Hubbard via Jordan-Wigner L = 10, d = 4, complex128): every local-update energy
    within 1e-10 relative of the unmodified reference, same bond dimensions and discarded weights
    (tests/golden/dmrg_c45.npz, generator make_golden_c45.py)."""
    from quantum_simulation_b200 import svd as qsvd
    if split == "gram":
        monkeypatch.setattr(qsvd, "FAST_MIN", 2)
    z = load_golden("dmrg_c45")
    L, d, chi0, seed, ns = (int(x) for x in z[f"{name}/cfg"])
    dt = torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64
    mps = quiet(qsb.MPS, L, d, chi0, "cpu", dt, seed)
    sums = np.array([t.detach().abs().sum().item() for t in mps.B])
    assert np.allclose(sums, z[f"{name}/init_abs_sums"], rtol=1e-13, atol=0)
    mps = mps.to(DEV)
    sw = qsb.Sweeps(ns)
    sw.set_schedule(maxdim=[int(x) for x in z[f"{name}/maxdim"]], krylov_dim=[int(x) for x in z[f"{name}/krylov_dim"]],
                    cutoff=[float(x) for x in z[f"{name}/cutoff"]], reortho=[bool(x) for x in z[f"{name}/reortho"]],
                    noise=[0.0] * ns)
    dm = quiet(qsb.DMRG, mps, _c45_mpo(z, name, dt), device=DEV, dtype=dt)
    E, _ = quiet(dm, sw, dispon=0, maxit=2)
    E, Eg = np.array(E), z[f"{name}/energies"]
    assert len(E) == len(Eg)
    assert dm.history["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    rel = np.max(np.abs(E - Eg) / np.abs(Eg))
    assert rel < 1e-10, f"max relative energy deviation {rel:.2e}"
    if split == "default":
        assert np.allclose(np.array(dm.history["discarded_weights"]), z[f"{name}/discarded"], rtol=1e-6, atol=1e-12)
    mpo = _c45_mpo(z, name, dt)
    assert abs(float(qsb.energy_expectation(dm.mps, mpo)) - E[-1]) < 1e-9

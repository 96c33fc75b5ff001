"""End-to-end GPU parity: DMRG(...)(sweeps) through our drop-in classes against the per-update energies,
bond dimensions and discarded weights recorded from the unmodified reference (tests/golden/dmrg.npz).
Tolerance (BASELINE.json): energy within 1e-10 relative, checked at EVERY local update, not just per sweep."""
import io
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

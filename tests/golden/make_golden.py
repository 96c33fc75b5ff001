#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the reference tree does not travel to the GPU
box):  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

What is written (all produced by importing quantum_simulation itself):
  mpo.npz       MPO site tensors of every model the parity tests / bench use
                (stock builders + FSM.add_term assemblies of SURVEY.md App. B)
  matvec.npz    ApplyMPO('ncon') inputs + outputs on small seeded cases
  env.npz       LeftEnvUpdate('ncon') / RightEnvUpdate('ncon') inputs + outputs
  lanczos.npz   lanczos_ground_state energies on seeded Hermitian matrices
  dmrg.npz      DMRG(...)(sweeps) per-update energies / bond dims / discarded
                weights for the notebook cases and BASELINE config 1
The script also asserts that oracle/dmrg_oracle.py reproduces each value, so a
successful run is the oracle's pin.
"""
import os
import sys
import types
import io
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for name in ("matplotlib", "matplotlib.pyplot", "IPython", "IPython.display"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["IPython"].display = sys.modules["IPython.display"]
sys.modules["IPython.display"].display = lambda *a, **k: None
sys.modules["IPython.display"].Math = lambda *a, **k: None
sys.path.insert(0, "/root/reference")
import warnings
warnings.filterwarnings("ignore")
import quantum_simulation as qs  # noqa: E402
from quantum_simulation.dmrg import ApplyMPO, LeftEnvUpdate, RightEnvUpdate, DMRG, Sweeps  # noqa: E402
from quantum_simulation.mps import MPS  # noqa: E402
from quantum_simulation.eigensolver import lanczos_ground_state as ref_lanczos  # noqa: E402
from quantum_simulation.mpo import (FSM, NamedData, Heisenberg1DMPOBuilder, Ising1DMPOBuilder,  # noqa: E402
                                    Ising2DMPOBuilder, generate_mpo_spin_operators)

from oracle import dmrg_oracle as orc  # noqa: E402

torch.set_num_threads(8)
C128, F64 = torch.complex128, torch.float64


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------ MPOs
def fsm_to_torch(fsm, dtype):
    out = []
    for m in fsm.to_mpo():
        t = torch.from_numpy(np.asarray(m))
        if not dtype.is_complex:
            assert abs(t.imag).max() == 0 if t.is_complex() else True
            t = t.real if t.is_complex() else t
        out.append(t.to(dtype).contiguous())
    return out


def heisenberg_real_mpo(L, J=1.0, Jz=None):
    """Real-valued spin-1/2 XXZ chain via FSM with Sp/Sm/Sz."""
    ops = generate_mpo_spin_operators(2)
    Sp, Sm, Sz = NamedData('Sp', ops['Sp']), NamedData('Sm', ops['Sm']), NamedData('Sz', ops['Sz'])
    Jz = J if Jz is None else Jz
    fsm = FSM(site_num=L)
    for i in range(L - 1):
        fsm.add_term(Jz, [Sz, Sz], [i, i + 1])
        fsm.add_term(J / 2, [Sp, Sm], [i, i + 1])
        fsm.add_term(J / 2, [Sm, Sp], [i, i + 1])
    return fsm_to_torch(fsm, F64)


def j1j2_cylinder_mpo(Lx, Ly=6, J1=1.0, J2=0.5):
    """J1-J2 Heisenberg cylinder, periodic in y, site = x*Ly + y."""
    ops = generate_mpo_spin_operators(2)
    Sp, Sm, Sz = NamedData('Sp', ops['Sp']), NamedData('Sm', ops['Sm']), NamedData('Sz', ops['Sz'])
    fsm = FSM(site_num=Lx * Ly)

    def bond(i, j, J):
        i, j = min(i, j), max(i, j)
        fsm.add_term(J, [Sz, Sz], [i, j])
        fsm.add_term(J / 2, [Sp, Sm], [i, j])
        fsm.add_term(J / 2, [Sm, Sp], [i, j])

    for x in range(Lx):
        for y in range(Ly):
            i = x * Ly + y
            bond(i, x * Ly + (y + 1) % Ly, J1)
            if x < Lx - 1:
                bond(i, (x + 1) * Ly + y, J1)
                bond(i, (x + 1) * Ly + (y + 1) % Ly, J2)
                bond(i, (x + 1) * Ly + (y - 1) % Ly, J2)
    return fsm_to_torch(fsm, F64)


def hubbard_jw_mpo(L, t=1.0, U=4.0):
    """Fermi-Hubbard chain, d=4 basis {0, up, dn, updn}, JW strings absorbed."""
    I2 = np.eye(2)
    a = np.array([[0., 1.], [0., 0.]])          # annihilate on one mode (|0>,|1>)
    Z = np.diag([1., -1.])
    cu = np.kron(a, I2)                          # c_up   (mode order: up, dn)
    cd = np.kron(Z, a)                           # c_dn with on-site string over up
    F = np.kron(Z, Z)                            # site parity
    nu, nd = cu.T @ cu, cd.T @ cd
    N = lambda n, m: NamedData(n, m)
    fsm = FSM(site_num=L)
    for i in range(L):
        fsm.add_term(U, [N('nund', nu @ nd)], [i])
    for i in range(L - 1):
        # c^dag_i c_{i+1} = (c^dag_i F_i) c_{i+1};  h.c. = (F_i c_i)^... written explicitly
        fsm.add_term(-t, [N('cudF', cu.T @ F), N('cu', cu)], [i, i + 1])
        fsm.add_term(-t, [N('cddF', cd.T @ F), N('cd', cd)], [i, i + 1])
        fsm.add_term(-t, [N('Fcu', F @ cu), N('cud', cu.T)], [i, i + 1])
        fsm.add_term(-t, [N('Fcd', F @ cd), N('cdd', cd.T)], [i, i + 1])
    return fsm_to_torch(fsm, F64)


def pack_mpo(store, name, mpo):
    """Store an MPO as its distinct site tensors plus an index list."""
    uniq, index = [], []
    for m in mpo:
        a = m.detach().cpu().numpy()
        for u, b in enumerate(uniq):
            if b.shape == a.shape and np.array_equal(a, b):
                index.append(u)
                break
        else:
            uniq.append(a)
            index.append(len(uniq) - 1)
    store[f"{name}/index"] = np.asarray(index, dtype=np.int64)
    for u, a in enumerate(uniq):
        store[f"{name}/t{u}"] = a


def gen_mpos():
    store = {}
    models = {}
    models["heis_c128_L6"] = quiet(Heisenberg1DMPOBuilder, 6, 1.0, 1.0, 1.0, dtype=C128).get_mpo()
    models["heis_c128_L4"] = quiet(Heisenberg1DMPOBuilder, 4, 1.0, 1.0, 1.0, dtype=C128).get_mpo()
    models["heis_c128_L20"] = quiet(Heisenberg1DMPOBuilder, 20, 1.0, 1.0, 1.0, dtype=C128).get_mpo()
    models["xxz_c128_L20"] = quiet(Heisenberg1DMPOBuilder, 20, 1.0, 1.0, 0.5, dtype=C128).get_mpo()
    models["heis_c128_L100"] = quiet(Heisenberg1DMPOBuilder, 100, 1.0, 1.0, 1.0, dtype=C128).get_mpo()
    models["heis_f64_L20"] = quiet(heisenberg_real_mpo, 20)
    models["xxz_f64_L20"] = quiet(heisenberg_real_mpo, 20, 1.0, 0.5)
    models["heis_f64_L100"] = quiet(heisenberg_real_mpo, 100)
    models["tfim_c128_L6_g0.7"] = quiet(Ising1DMPOBuilder, 6, 1.0, 0.7, dtype=C128).get_mpo()
    models["tfim_c128_L6_g0.75"] = quiet(Ising1DMPOBuilder, 6, 1.0, 0.75, dtype=C128).get_mpo()
    models["tfim_c128_L40_g0.5"] = quiet(Ising1DMPOBuilder, 40, 1.0, 0.5, dtype=C128).get_mpo()
    models["tfim_c128_L100_g0.5"] = quiet(Ising1DMPOBuilder, 100, 1.0, 0.5, dtype=C128).get_mpo()
    models["ising2d_c128_2x2"] = quiet(Ising2DMPOBuilder, 2, 2, 1.0, 0.5, dtype=C128).get_mpo()
    models["j1j2_f64_8x6"] = quiet(j1j2_cylinder_mpo, 8, 6)
    models["hubbard_f64_L64"] = quiet(hubbard_jw_mpo, 64)
    for k, v in models.items():
        v = [t.detach() for t in v]
        models[k] = v
        pack_mpo(store, k, v)
        print(f"  mpo {k}: bonds {[tuple(t.shape[:2]) for t in v[:4]]} ... d={v[0].shape[2]}")
    np.savez_compressed(os.path.join(HERE, "mpo.npz"), **store)
    return models


# ------------------------------------------------------------------ matvec / env
def randn(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, dtype=dtype, generator=g)


def gen_matvec(models):
    ref_ncon, ref_ein = ApplyMPO('ncon'), ApplyMPO('einsum')
    store, names = {}, []
    hb = models["heis_c128_L20"]
    hr = models["heis_f64_L20"]
    tf = models["tfim_c128_L40_g0.5"]
    hub = models["hubbard_f64_L64"]
    cyl = models["j1j2_f64_8x6"]
    rnd3 = [randn((3, 4, 3, 3), F64, 90), randn((4, 2, 3, 3), F64, 91)]
    # (name, M1, M2, chi_l_out, chi_l_in, chi_r_out, chi_r_in, dtype, layout)
    cases = [
        ("edge_first_c128", hb[0], hb[1], 1, 1, 4, 4, C128, "contig"),
        ("edge_second_c128", hb[1], hb[2], 2, 2, 6, 6, C128, "contig"),
        ("bulk6_c128", hb[5], hb[6], 6, 6, 6, 6, C128, "contig"),
        ("edge_last_c128", hb[18], hb[19], 4, 4, 1, 1, C128, "contig"),
        ("bulk6_f64", hr[5], hr[6], 6, 6, 6, 6, F64, "contig"),
        ("edge_first_f64", hr[0], hr[1], 1, 1, 4, 4, F64, "contig"),
        ("edge_last_f64", hr[18], hr[19], 4, 4, 1, 1, F64, "contig"),
        ("odd17_f64", hr[5], hr[6], 17, 17, 5, 5, F64, "contig"),
        ("odd17_c128", hb[5], hb[6], 17, 17, 5, 5, C128, "ncon"),
        ("rect_f64", hr[5], hr[6], 9, 12, 7, 10, F64, "contig"),
        ("rect_c128", hb[5], hb[6], 9, 12, 7, 10, C128, "contig"),
        ("d3_rand_f64", rnd3[0], rnd3[1], 8, 8, 5, 5, F64, "contig"),
        ("d3_rand_c128", rnd3[0].to(C128) * (1 + 0.5j), rnd3[1].to(C128) * (0.3 - 1j), 8, 8, 5, 5, C128, "ncon"),
        ("hubbard_d4_f64", hub[10], hub[11], 12, 12, 12, 12, F64, "contig"),
        ("hubbard_d4_c128", hub[10].to(C128), hub[11].to(C128), 12, 12, 10, 10, C128, "ncon"),
        ("tfim16_c128", tf[10], tf[11], 16, 16, 16, 16, C128, "contig"),
        ("cyl_f64", cyl[20], cyl[21], 16, 16, 12, 12, F64, "ncon"),
        ("bulk32_f64_ncon", hr[5], hr[6], 32, 32, 32, 32, F64, "ncon"),
        ("bulk64_f64", hr[5], hr[6], 64, 64, 64, 64, F64, "contig"),
        ("bulk64_c128", hb[5], hb[6], 64, 64, 64, 64, C128, "ncon"),
        ("bulk40x24_c128", hb[5], hb[6], 40, 40, 24, 24, C128, "contig"),
    ]
    for ci, (name, M1, M2, clo, cli, cro, cri, dt, layout) in enumerate(cases):
        M1, M2 = M1.to(dt), M2.to(dt)
        Wl, Wr = M1.shape[0], M2.shape[1]
        L = randn((Wl, clo, cli), dt, 1000 + ci)
        R = randn((Wr, cro, cri), dt, 2000 + ci)
        if layout == "ncon":   # the strided layout the reference's own env updates emit: strides (chi, 1, W*chi)
            L = L.permute(2, 0, 1).contiguous().permute(1, 2, 0)
            R = R.permute(2, 0, 1).contiguous().permute(1, 2, 0)
        psi = randn((cli * M1.shape[3] * M2.shape[3] * cri,), dt, 3000 + ci)
        psi = psi / torch.linalg.norm(psi)
        out = ref_ncon(psi, L, M1, M2, R).clone()
        out2 = ref_ein(psi, L, M1, M2, R)
        assert torch.linalg.norm(out - out2) <= 1e-13 * torch.linalg.norm(out), name
        o = orc.apply_mpo(psi, L, M1, M2, R)
        err = (torch.linalg.norm(o - out) / torch.linalg.norm(out)).item()
        assert err == 0.0, (name, err)          # same tensordots on the same MKL: bit-identical
        buf = torch.empty_like(out)
        assert ref_ncon(psi, L, M1, M2, R, out=buf).data_ptr() == buf.data_ptr()
        assert torch.equal(buf, out)
        names.append(name)
        for k, v in dict(psi=psi, L=L.contiguous(), M1=M1, M2=M2, R=R.contiguous(), out=out).items():
            store[f"{name}/{k}"] = v.numpy()
        store[f"{name}/layout"] = np.array(layout)
        print(f"  matvec {name}: n={psi.numel()} ok")
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "matvec.npz"), **store)


def gen_env(models):
    lup, rup = LeftEnvUpdate('ncon'), RightEnvUpdate('ncon')
    lup_e, rup_e = LeftEnvUpdate('einsum'), RightEnvUpdate('einsum')
    store, names = {}, []
    hb, hr = models["heis_c128_L20"], models["heis_f64_L20"]
    hub, cyl = models["hubbard_f64_L64"], models["j1j2_f64_8x6"]
    # (name, M, chi, chi', dtype, mps-tensor layout)
    cases = [
        ("first_c128", hb[0], 1, 2, C128, "contig"),
        ("second_c128", hb[1], 2, 4, C128, "svd"),
        ("bulk_c128", hb[5], 6, 6, C128, "svd"),
        ("shrink_c128", hb[17], 8, 4, C128, "contig"),
        ("last_c128", hb[19], 2, 1, C128, "contig"),
        ("bulk_f64", hr[5], 6, 6, F64, "svd"),
        ("odd_f64", hr[5], 17, 13, F64, "svd"),
        ("grow_f64", hr[5], 16, 32, F64, "contig"),
        ("hubbard_f64", hub[10], 12, 20, F64, "svd"),
        ("hubbard_c128", hub[10].to(C128), 9, 12, C128, "svd"),
        ("cyl_f64", cyl[20], 16, 24, F64, "contig"),
        ("bulk48_c128", hb[5], 48, 48, C128, "svd"),
        ("bulk64_f64", hr[5], 64, 64, F64, "svd"),
    ]
    for ci, (name, M, chi, chip, dt, layout) in enumerate(cases):
        M = M.to(dt)
        W, Wp, d = M.shape[0], M.shape[1], M.shape[2]
        # left update: Lp (W, chi, chi), A (chi, d, chip)  -> (Wp, chip, chip)
        Lp = randn((W, chi, chi), dt, 4000 + ci)
        A = randn((chi, d, chip), dt, 5000 + ci)
        # right update: Rn (Wp, chip, chip), B (chi, d, chip) -> (W, chi, chi)
        Rn = randn((Wp, chip, chip), dt, 6000 + ci)
        B = randn((chi, d, chip), dt, 7000 + ci)
        if layout == "svd":
            # the strides DMRG hands over: A = U[:, :k].view(...) column-major,
            # B = Vh[:k, :].view(...) column-major and row-truncated
            A = A.reshape(chi * d, chip).T.contiguous().T.view(chi, d, chip)
            big = torch.zeros(chi + 3, d * chip, dtype=dt)
            big[:chi] = B.reshape(chi, d * chip)
            B = big.T.contiguous().T[:chi, :].view(chi, d, chip)
        Ln = lup(Lp, M, A)
        Rp = rup(M, Rn, B)
        assert tuple(Ln.shape) == (Wp, chip, chip) and tuple(Rp.shape) == (W, chi, chi)
        assert torch.linalg.norm(Ln - lup_e(Lp, M, A)) <= 1e-13 * torch.linalg.norm(Ln)
        assert torch.linalg.norm(Rp - rup_e(M, Rn, B)) <= 1e-13 * torch.linalg.norm(Rp)
        oL, oR = orc.left_env_update(Lp, M, A), orc.right_env_update(M, Rn, B)
        assert torch.equal(oL, Ln) and torch.equal(oR, Rp), name
        assert oL.stride() == Ln.stride() and oR.stride() == Rp.stride()
        names.append(name)
        for k, v in dict(Lp=Lp, M=M, A=A.contiguous(), Rn=Rn, B=B.contiguous(),
                         Lnew=Ln.contiguous(), Rnew=Rp.contiguous()).items():
            store[f"{name}/{k}"] = v.numpy()
        store[f"{name}/layout"] = np.array(layout)
        print(f"  env {name}: ok, ncon strides {Ln.stride()}")
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "env.npz"), **store)


# ------------------------------------------------------------------ lanczos
def rand_herm(n, dtype, seed):
    A = randn((n, n), dtype, seed)
    H = 0.5 * (A + A.conj().T)
    return H + 0.1 * torch.eye(n, dtype=dtype)


def gen_lanczos():
    store, names = {}, []
    cases = [("f64_n32_k12_pro", 32, F64, 12, 6, True, 1), ("c128_n24_k8", 24, C128, 8, 4, False, 1),
             ("f64_n64_k4", 64, F64, 4, 2, False, 1), ("c128_n40_k6_pro2", 40, C128, 6, 3, True, 2)]
    mv = lambda x, A: A @ x
    for ci, (name, n, dt, k, r, pro, mr) in enumerate(cases):
        H = rand_herm(n, dt, 100 + ci)
        v0 = randn((n,), dt, 200 + ci)
        v, E = ref_lanczos(v0, mv, (H,), num_restarts=r, krylov_dim=k, pro=pro, pro_max_reorth=mr)
        v = v.clone()
        vo, Eo = orc.lanczos_ground_state(v0, mv, (H,), num_restarts=r, krylov_dim=k, pro=pro, pro_max_reorth=mr)
        assert E == Eo and torch.equal(v, vo), name
        names.append(name)
        store[f"{name}/H"], store[f"{name}/v0"], store[f"{name}/v"] = H.numpy(), v0.numpy(), v.numpy()
        store[f"{name}/E"] = np.array(E)
        store[f"{name}/E_exact"] = np.array(torch.linalg.eigvalsh(H)[0].item())
        store[f"{name}/params"] = np.array([k, r, int(pro), mr])
        print(f"  lanczos {name}: E={E:.12f}")
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "lanczos.npz"), **store)


# ------------------------------------------------------------------ full DMRG runs
def run_ref_dmrg(mpo, L, d, chi0, dtype, seed, sched, backend='ncon', maxit=2):
    mps = quiet(MPS, L, d, chi0, 'cpu', dtype, seed)
    init_sum = np.array([t.detach().abs().sum().item() for t in mps.B])
    sw = Sweeps(sched['numsweeps'])
    sw.set_schedule(**{k: v for k, v in sched.items() if k != 'numsweeps'})
    dm = quiet(DMRG, mps, mpo, device='cpu', dtype=dtype, contract_backend=backend)
    E, _ = quiet(dm, sw, dispon=0, maxit=maxit)
    return np.array(E), dm.history, init_sum


def run_orc_dmrg(mpo, L, d, chi0, dtype, seed, sched, maxit=2):
    mps = orc.OracleMPS(L, d, chi0, dtype, seed)
    init_sum = np.array([t.abs().sum().item() for t in mps.B])
    E, h = orc.dmrg_run(mps, mpo, numsweeps=sched['numsweeps'], maxdim=sched['maxdim'],
                        krylov_dim=sched['krylov_dim'], cutoff=sched['cutoff'],
                        reortho=sched['reortho'], maxreortho=sched.get('maxreortho', [1] * sched['numsweeps']),
                        maxit=maxit)
    return np.array(E), h, init_sum


def gen_dmrg(models):
    store, names = {}, []
    n3 = dict(numsweeps=3, maxdim=[4, 8, 12], krylov_dim=[4, 6, 8], cutoff=[1e-8, 1e-10, 1e-12],
              reortho=[False, False, True], noise=[0.0] * 3)
    cases = [
        # notebook dmrg_02 (noise set to 0: the notebook's noise draws from the global RNG)
        ("nb02_tfim_L6", "tfim_c128_L6_g0.7", 6, 2, 4, C128, 2024, n3, 'ncon', -2.327594419162),
        # notebook dmrg_03: Heisenberg L=4, seed 7, 1 sweep maxdim 8, krylov 6, reortho
        ("nb03_heis_L4", "heis_c128_L4", 4, 2, 8, C128, 7,
         dict(numsweeps=1, maxdim=[8], krylov_dim=[6], cutoff=[1e-12], reortho=[True], noise=[0.0]), 'ncon',
         -1.616025403784),
        ("nb03_ising2d", "ising2d_c128_2x2", 4, 2, 8, C128, 11,
         dict(numsweeps=1, maxdim=[8], krylov_dim=[6], cutoff=[1e-12], reortho=[True], noise=[0.0]), 'ncon',
         -1.306562964876),
        # SURVEY section 4 calibration runs: maxdim 32 x3, cutoff 0, krylov 4, seed 1
        ("cal_heis_L20", "heis_c128_L20", 20, 2, 32, C128, 1,
         dict(numsweeps=3, maxdim=[32] * 3, krylov_dim=[4] * 3, cutoff=[0.0] * 3, reortho=[False] * 3,
              noise=[0.0] * 3), 'ncon', -8.682473319650),
        ("cal_tfim_L40", "tfim_c128_L40_g0.5", 40, 2, 32, C128, 1,
         dict(numsweeps=3, maxdim=[32] * 3, krylov_dim=[4] * 3, cutoff=[0.0] * 3, reortho=[False] * 3,
              noise=[0.0] * 3), 'ncon', -12.642358448079),
        # BASELINE config 1: XXZ L=20 chi=64 (complex, stock builder) and the real-MPO variant
        ("c1_xxz_L20_chi64_c128", "xxz_c128_L20", 20, 2, 64, C128, 1,
         dict(numsweeps=2, maxdim=[64] * 2, krylov_dim=[4] * 2, cutoff=[0.0] * 2, reortho=[False] * 2,
              noise=[0.0] * 2), 'ncon', None),
        ("c1_xxz_L20_chi64_f64", "xxz_f64_L20", 20, 2, 64, F64, 1,
         dict(numsweeps=2, maxdim=[64] * 2, krylov_dim=[4] * 2, cutoff=[0.0] * 2, reortho=[False] * 2,
              noise=[0.0] * 2), 'ncon', None),
        ("heis_f64_L20_pro", "heis_f64_L20", 20, 2, 16, F64, 3,
         dict(numsweeps=2, maxdim=[16, 24], krylov_dim=[5, 6], cutoff=[1e-10, 1e-12], reortho=[True, True],
              maxreortho=[1, 2], noise=[0.0] * 2), 'ncon', None),
    ]
    for name, mk, L, d, chi0, dt, seed, sched, be, pinned in cases:
        mpo = [m.to(dt) for m in models[mk]]
        E, h, s0 = run_ref_dmrg(mpo, L, d, chi0, dt, seed, sched, be)
        osched = dict(sched)
        Eo, ho, s0o = run_orc_dmrg(mpo, L, d, chi0, dt, seed, osched)
        assert np.array_equal(s0, s0o), f"{name}: oracle MPS init differs"
        assert ho['bond_dims'] == h['bond_dims'], name
        rel = np.max(np.abs(E - Eo) / np.abs(E))
        assert rel < 1e-13, (name, rel)
        if pinned is not None:
            assert abs(E[-1] - pinned) < 5e-12 * max(1, abs(pinned)) + 5e-13, (name, E[-1], pinned)
        names.append(name)
        store[f"{name}/energies"] = E
        store[f"{name}/bond_dims"] = np.array(h['bond_dims'])
        store[f"{name}/discarded"] = np.array(h['discarded_weights'])
        store[f"{name}/init_abs_sums"] = s0
        store[f"{name}/mpo"] = np.array(mk)
        store[f"{name}/cfg"] = np.array([L, d, chi0, seed, sched['numsweeps']])
        store[f"{name}/dtype"] = np.array("c128" if dt == C128 else "f64")
        for k in ('maxdim', 'krylov_dim', 'cutoff', 'reortho'):
            store[f"{name}/{k}"] = np.array(sched[k])
        store[f"{name}/maxreortho"] = np.array(sched.get('maxreortho', [1] * sched['numsweeps']))
        print(f"  dmrg {name}: E_final={E[-1]:.12f}  updates={len(E)}  oracle rel diff={rel:.1e}")
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "dmrg.npz"), **store)


if __name__ == "__main__":
    print("MPOs");     models = gen_mpos()
    print("matvec");   gen_matvec(models)
    print("env");      gen_env(models)
    print("lanczos");  gen_lanczos()
    print("dmrg");     gen_dmrg(models)
    print("done")

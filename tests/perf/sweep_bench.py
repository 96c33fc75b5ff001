#!/usr/bin/env python3
"""Sweep-level measurements (BASELINE.json: "DMRG sweep time"): full DMRG runs through the drop-in DMRG class on
the GPU, energy parity against the CPU oracle where it finishes in seconds, and a per-bond-step breakdown at
chi = 2048.  Prints one JSON object per case.  Usage: python tests/perf/sweep_bench.py [c1] [c2] [c3] [svd]"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantum_simulation_b200 as qsb          # noqa: E402
from oracle import dmrg_oracle as orc          # noqa: E402

DEV = "cuda"


def mpo(key, dtype):
    z = np.load(os.path.join(ROOT, "tests", "golden", "mpo.npz"))
    return [torch.from_numpy(z[f"{key}/t{int(u)}"]).to(dtype).contiguous() for u in z[f"{key}/index"]]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_gpu(key, L, chi, dtype, nsweeps, seed=1, krylov=4, maxit=2):
    mps = quiet(qsb.MPS, L, 2, chi, "cpu", dtype, seed).to(DEV)
    sw = qsb.Sweeps(nsweeps)
    sw.set_schedule(maxdim=[chi] * nsweeps, cutoff=[0.0] * nsweeps, krylov_dim=[krylov] * nsweeps,
                    noise=[0.0] * nsweeps)
    dm = quiet(qsb.DMRG, mps, [m.to(DEV) for m in mpo(key, dtype)], device=DEV, dtype=dtype)
    dm.profile = "--profile" in sys.argv
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E, _ = quiet(dm, sw, dispon=0, maxit=maxit)
    torch.cuda.synchronize()
    return np.array(E), time.perf_counter() - t0, dm


def run_cpu(key, L, chi, dtype, nsweeps, seed=1, krylov=4, maxit=2):
    torch.set_num_threads(os.cpu_count() or 1)
    mps = orc.OracleMPS(L, 2, chi, dtype, seed)
    t0 = time.perf_counter()
    E, h = orc.dmrg_run(mps, mpo(key, dtype), numsweeps=nsweeps, maxdim=[chi] * nsweeps, krylov_dim=[krylov] * nsweeps,
                        cutoff=[0.0] * nsweeps, reortho=[False] * nsweeps, maxreortho=[1] * nsweeps, maxit=maxit)
    return np.array(E), time.perf_counter() - t0


def case_c1():
    """BASELINE config 1: XXZ chain L=20, chi=64, full run on both sides."""
    out = []
    for key, dt in (("xxz_c128_L20", torch.complex128), ("xxz_f64_L20", torch.float64)):
        run_gpu(key, 20, 64, dt, 1)                       # warm-up (workspaces, cuSOLVER handles)
        Eg, tg, dm = run_gpu(key, 20, 64, dt, 3)
        Ec, tc = run_cpu(key, 20, 64, dt, 3)
        nb = len(Eg) // 3
        rel = np.abs(Eg - Ec) / np.abs(Ec)
        out.append({"case": "C1 XXZ L=20 chi=64, 3 sweeps", "mpo": key, "gpu_s": tg, "cpu_oracle_s": tc,
                    "cpu_cores": os.cpu_count(), "gpu_s_per_sweep": tg / 3, "cpu_s_per_sweep": tc / 3,
                    "E_final_gpu": float(Eg[-1]), "E_final_cpu": float(Ec[-1]),
                    "max_rel_dev_all_updates": float(rel.max()),
                    "rel_dev_end_of_sweep": [float(rel[(k + 1) * nb - 1]) for k in range(3)]})
    return out


def case_c2():
    """BASELINE config 2: TFIM L=100, chi=512, 10 sweeps on the GPU; 1 oracle sweep on the CPU for the ratio."""
    key, dt = "tfim_c128_L100_g0.5", torch.complex128
    if "--first-sweep" in sys.argv:         # per-update energies of sweep 1, both two-site-split paths
        from quantum_simulation_b200 import svd as qsvd
        out = []
        for fm, tag in ((256, "gram"), (10 ** 9, "torchsvd")):
            qsvd.FAST_MIN = fm
            Eg, tg, dm = run_gpu(key, 100, 512, dt, 1)
            np.save(os.path.join(ROOT, "gpurun_out", f"c2_first_sweep_{tag}.npy"), Eg)
            out.append({"case": f"C2 first sweep ({tag})", "gpu_s": tg, "E_end": float(Eg[-1])})
        qsvd.FAST_MIN = 256
        return out
    Eg, tg, dm = run_gpu(key, 100, 512, dt, 10)
    t0 = time.perf_counter()
    mp = [m.to(DEV) for m in mpo(key, dt)]
    e_env = float(qsb.energy_expectation(dm.mps, mp))
    var_env = float(qsb.energy_variance(dm.mps, mp))
    torch.cuda.synchronize()
    t_obs = time.perf_counter() - t0
    res = {"case": "C2 TFIM L=100 g=0.5 chi=512, 10 sweeps", "mpo": key, "gpu_s": tg, "gpu_s_per_sweep": tg / 10,
           "phase_seconds": dict(dm.timers) if dm.profile else None,
           "split_paths": {k_: v_ for k_, v_ in __import__("quantum_simulation_b200").svd.stats.items() if not isinstance(v_, list)},
           "launches": qsb.launch_counters(),
           "energy_by_environments": e_env, "variance_by_environments": var_env, "observables_s": t_obs,
           "E_final_gpu": float(Eg[-1]), "bond_dim_mid": dm.history["bond_dims"][-50],
           "E_end_of_sweep": [float(Eg[(k + 1) * 198 - 1]) for k in range(10)]}
    if "--cpu" in sys.argv:
        Ec, tc = run_cpu(key, 100, 512, dt, 1)
        res.update({"cpu_oracle_s_first_sweep": tc, "cpu_cores": os.cpu_count(),
                    "max_rel_dev_first_sweep": float((np.abs(Eg[:198] - Ec) / np.abs(Ec)).max())})
    return [res]


def time_it(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def case_svd():
    """The non-hot-path cost that dominates a chi=2048 sweep: SVD of theta (chi*d x d*chi)."""
    out = []
    for n in (1024, 2048, 4096):
        a = torch.randn(n, n, dtype=torch.float64, device=DEV)
        row = {"case": f"svd {n}x{n} f64"}
        for drv in (None, "gesvdj", "gesvda"):
            try:
                row[str(drv)] = time_it(lambda: torch.linalg.svd(a, full_matrices=False, driver=drv), 1)
            except Exception as ex:
                row[str(drv)] = repr(ex)[:80]
        row["eigh_gram"] = time_it(lambda: torch.linalg.eigh(a @ a.T), 1)
        row["qr"] = time_it(lambda: torch.linalg.qr(a), 1)
        out.append(row)
    return out


def case_c3():
    """Per-bond-step breakdown at BASELINE config 3 (Heisenberg L=100, chi=2048, f64): one bulk bond step."""
    import bench
    chi, dt = 2048, torch.float64
    L, M1, M2, R, theta = bench.make_inputs(chi, dt, torch.device(DEV))
    L = L + L.transpose(1, 2)
    R = R + R.transpose(1, 2)
    op = qsb.ApplyMPO()
    t_l = time_it(lambda: qsb.lanczos_ground_state(theta, op, (L, M1, M2, R), num_restarts=2, krylov_dim=4), 2)
    A = torch.randn(chi, 2, chi, dtype=dt, device=DEV)
    t_e = time_it(lambda: qsb.LeftEnvUpdate()(L, M1, A), 2)
    t_th = time_it(lambda: torch.matmul(A.reshape(chi * 2, chi), A.reshape(chi, 2 * chi)), 2)
    th = theta.view(chi * 2, 2 * chi)
    t_svd = time_it(lambda: torch.linalg.svd(th, full_matrices=False), 1)
    step = t_l + t_e + t_th + t_svd
    return [{"case": "C3 Heisenberg L=100 chi=2048 f64: one bulk bond step", "lanczos_8_matvecs_s": t_l,
             "env_update_s": t_e, "theta_build_s": t_th, "svd_torch_default_s": t_svd, "bond_step_s": step,
             "sweep_198_steps_estimate_s": 198 * step, "hot_path_share": (t_l + t_e) / step}]


def case_c3full():
    """BASELINE config 3, ONE FULL SWEEP measured (not estimated): Heisenberg L=100, chi=2048, float64, MPS built and
    canonicalised on the GPU, krylov_dim=4, maxit=2, cutoff=0 -- 198 bond updates."""
    dt = torch.float64
    t0 = time.perf_counter()
    mps = quiet(qsb.MPS, 100, 2, 2048, DEV, dt, 1)
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    sw = qsb.Sweeps(1)
    sw.set_schedule(maxdim=[2048], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
    dm = quiet(qsb.DMRG, mps, [m.to(DEV) for m in mpo("heis_f64_L100", dt)], device=DEV, dtype=dt)
    import quantum_simulation_b200.svd as _qsvd
    _qsvd.TIMING = "--split-timing" in sys.argv
    dm.profile = "--split-timing" in sys.argv
    torch.manual_seed(1)       # zero centre tensor at this size -> first Lanczos start is a global-RNG draw
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E, _ = quiet(dm, sw, dispon=0, maxit=2)
    torch.cuda.synchronize()
    t_sweep = time.perf_counter() - t0
    bd = dm.history["bond_dims"]
    return [{"case": "C3 Heisenberg L=100 chi=2048 f64: one full sweep (198 bond updates), measured",
             "mps_init_on_gpu_s": t_init, "sweep_s": t_sweep, "E_end_of_sweep": float(E[-1]),
             "E_mid_right": float(E[98]), "bond_dim_max": int(max(bd)), "updates_at_full_chi": int(sum(b == 2048 for b in bd)),
             "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9,
             "phase_seconds": dict(dm.timers) if dm.profile else None,
             "split_paths": {k_: (v_ if not isinstance(v_, list) else {"n": len(v_), "min": min(v_), "median": sorted(v_)[len(v_) // 2], "max": max(v_), "first10": v_[:10], "last10": v_[-10:]})
                             for k_, v_ in __import__("quantum_simulation_b200").svd.stats.items()},
             "discarded_weight_max": float(max(dm.history["discarded_weights"])),
             "discarded_weight_max_full_chi": float(max(w for w, b in zip(dm.history["discarded_weights"], bd) if b == 2048))}]


def case_c3multi(nsweeps=3):
    """BASELINE config 3, several sweeps: the time and the end-of-sweep energy of every sweep (the first one starts from
    the random state, later ones are in the converged regime: every split takes the range-compressed path and the
    Lanczos solver stops where the reference stops)."""
    import quantum_simulation_b200.svd as _qsvd
    dt = torch.float64
    mps = quiet(qsb.MPS, 100, 2, 2048, DEV, dt, 1)
    mpo_ = [m.to(DEV) for m in mpo("heis_f64_L100", dt)]
    dm = quiet(qsb.DMRG, mps, mpo_, device=DEV, dtype=dt)
    torch.manual_seed(1)
    rows = []
    for k in range(nsweeps):
        sw = qsb.Sweeps(1)
        sw.set_schedule(maxdim=[2048], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
        before = dict(_qsvd.stats)
        lc0 = qsb.launch_counters()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        E, _ = quiet(dm, sw, dispon=0, maxit=2)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        lc1 = qsb.launch_counters()
        rows.append({"case": f"C3 Heisenberg L=100 chi=2048 f64, sweep {k + 1} of {nsweeps} on the same state", "sweep_s": t,
                     "E_end_of_sweep": float(E[-1]),
                     "split_paths": {key: _qsvd.stats[key] - before.get(key, 0) for key in
                                     ("sketch_accepted", "sketch_rejected", "sketch_skipped", "gram", "svd")},
                     "gemm_launches": lc1["tma_gemm"] - lc0["tma_gemm"],
                     "discarded_weight_max": float(max(dm.history["discarded_weights"]))})
    rows.append({"case": "C3 after the last sweep", "E_by_environments": float(qsb.energy_expectation(dm.mps, mpo_)),
                 "variance_by_environments": float(qsb.energy_variance(dm.mps, mpo_))})
    return rows


def case_split():
    """Stage timing of the two-site split at the C3 bulk shape (4096 x 4096 float64, keep 2048) on a block with a
    DMRG-like spectrum: the range-compressed path against the full Gram/eigh path."""
    from quantum_simulation_b200 import svd as qsvd
    dt = torch.float64
    n, k = 4096, 2048
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(n, n, dtype=dt, device=DEV, generator=g)
    U0 = torch.linalg.qr(a)[0]
    V0 = torch.linalg.qr(a.T.contiguous())[0]
    rows = []
    for decay in (60.0, 30.0):
        sv = torch.exp(-torch.arange(n, dtype=dt, device=DEV) * (decay / n) * 8)
        theta = (U0 * sv) @ V0.T
        theta = theta / torch.linalg.norm(theta)

        def tm(fn, reps=3):
            fn(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            return best * 1e3, r
        t_gram, (Ug, Sg, Vg, kg) = tm(lambda: qsvd._gram_split(theta, k, 0.0))
        qsvd._backoff.clear()
        t_sk, out = tm(lambda: qsvd._sketch_split(theta, k, 0.0))
        l = k + qsvd._oversample(k)
        om = qsvd._omega(n, l, dt, theta.device)
        t_y, Y = tm(lambda: qsvd._mm(theta, om))
        regs = {}
        for reg in (1e-9, 1e-10, 1e-11, 1e-12):
            Yr = Y + qsvd._fixed_normal("reg", n, l, dt, theta.device, 0xB200) * (reg * float(torch.linalg.norm(Y)) / (n * l) ** 0.5)
            for sp in (6,):
                t_, Qr = tm(lambda: qsvd._orthonormal_cols(Yr, sp), 1)
                if Qr is not None:
                    regs[f"reg{reg:g}_passes{sp}"] = {"ms": t_, "rho": float(torch.linalg.norm(theta - Qr @ (Qr.T @ theta)) ** 2),
                                                      "orth": float((Qr.T @ Qr - torch.eye(l, dtype=dt, device=DEV)).abs().max())}
                else:
                    regs[f"reg{reg:g}_passes{sp}"] = {"ms": t_, "failed": True}
        t_q, Q = tm(lambda: qsvd._orthonormal_cols(Y))
        t_hq, _ = tm(lambda: torch.linalg.qr(Y)[0])
        if Q is None:
            Q = torch.linalg.qr(Y)[0]
        t_b, B = tm(lambda: qsvd._mm(Q.T, theta))
        t_r, _ = tm(lambda: torch.linalg.norm(theta - qsvd._mm(Q, B)))
        Gs = qsvd._mm(B, B.T)
        t_e, _ = tm(lambda: torch.linalg.eigh(Gs))
        t_e4, _ = tm(lambda: torch.linalg.eigh(qsvd._mm(theta, theta.T)))
        Bk = B[:k].contiguous()
        t_o, _ = tm(lambda: qsvd._orthonormal_rows(Bk))
        row = {"case": f"two-site split 4096x4096 f64 keep 2048, S_j ~ exp(-{decay * 8 / n:.3f} j)", "gram_split_ms": t_gram,
               "stats": {k_: v_ for k_, v_ in qsvd.stats.items() if k_ != "rejected_rhos"},
               "sketch_split_ms": t_sk, "sketch_accepted": out is not None, "rho": qsvd.stats["last_rho"],
               "stages_ms": {"Y=theta*Omega": t_y, "sCholQR3(Y)": t_q, "cholqr_ok": qsvd._orthonormal_cols(Y) is not None,
                             "householder_qr(Y)": t_hq, "B=Q^H theta": t_b, "residual": t_r, f"eigh({l})": t_e,
                             "eigh(4096)+gram": t_e4, "orthonormal_rows(k x n)": t_o},
               "regularised_cholqr": regs}
        if out is not None:
            U, S, Vh, keep = out
            row["S2_maxdev"] = float((S[:k] ** 2 - Sg[:k] ** 2).abs().max())
            row["rec_dev"] = float(torch.linalg.norm((U * S[:k]) @ Vh - (Ug * Sg[:k]) @ Vg))
            row["iso_dev"] = [float(torch.linalg.norm(U.T @ U - torch.eye(k, dtype=dt, device=DEV))),
                              float(torch.linalg.norm(Vh @ Vh.T - torch.eye(k, dtype=dt, device=DEV)))]
        rows.append(row)
    return rows


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1"]
    for w in which:
        for row in {"c1": case_c1, "c2": case_c2, "c3": case_c3, "svd": case_svd, "c3full": case_c3full, "split": case_split, "c3multi": case_c3multi}[w]():
            print(json.dumps(row), flush=True)

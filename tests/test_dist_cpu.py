"""World-size-2 gloo tests (CPU) of the multi-GPU HOST logic: the row partition, the
all-gather exchange step of the sharded matvec and the scalar all-reduces of the sharded Lanczos.  The numerical
engine on each rank is the oracle (matvec) / a torch vector backend, injected through the same hooks the CUDA
path uses; the product defaults stay CUDA-only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(dtype, chi_l, chi_r, hermitian):
    from conftest import golden_mpo
    mpo = golden_mpo("heis_c128_L20" if dtype.is_complex else "heis_f64_L20", dtype)
    M1, M2 = mpo[9], mpo[10]
    g = torch.Generator().manual_seed(7)
    L = torch.randn(5, chi_l, chi_l, dtype=dtype, generator=g)
    R = torch.randn(5, chi_r, chi_r, dtype=dtype, generator=g)
    if hermitian:        # physical environments are Hermitian per MPO channel pair; make H_eff Hermitian
        L = L + L.conj().transpose(1, 2)
        R = R + R.conj().transpose(1, 2)
    psi = torch.randn(chi_l * 4 * chi_r, dtype=dtype, generator=g)
    return L, M1, M2, R, psi


def _worker(rank, world, port, chi_l, chi_r, dtype_name, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        import quantum_simulation_b200 as qsb
        from quantum_simulation_b200 import parallel
        from oracle import dmrg_oracle as orc
        from vecops_torch import TorchVectorOps
        dtype = getattr(torch, dtype_name)
        L, M1, M2, R, psi = _problem(dtype, chi_l, chi_r, hermitian=True)
        per_row = 4 * chi_r
        lo, hi = parallel.row_partition(chi_l, world)[rank]
        L_rows = parallel.shard_rows(L, 1, rank, world)
        assert L_rows.shape[1] == hi - lo
        op = parallel.ShardedApplyMPO(chi_l, per_row, local_op=lambda v, *a, out=None: orc.apply_mpo(v, *a, out=out))
        # 1. sharded matvec == rows of the full matvec
        full_ref = orc.apply_mpo(psi, L, M1, M2, R)
        shard_in = psi[lo * per_row: hi * per_row].clone()
        out = op(shard_in, L_rows, M1, M2, R)
        err_mv = float(torch.linalg.norm(out - full_ref[lo * per_row: hi * per_row]) / torch.linalg.norm(full_ref))
        gathered = parallel.gather_rows(out, chi_l, per_row)
        err_gather = float(torch.linalg.norm(gathered - full_ref) / torch.linalg.norm(full_ref))
        # 2. sharded Lanczos == unsharded oracle Lanczos
        _, E_ref = orc.lanczos_ground_state(psi, orc.apply_mpo, (L, M1, M2, R), num_restarts=2, krylov_dim=4)
        v_sh, E = parallel.lanczos_ground_state_sharded(shard_in, op, (L_rows, M1, M2, R), num_restarts=2,
                                                        krylov_dim=4, _vector_ops=TorchVectorOps())
        _, E_pro_ref = orc.lanczos_ground_state(psi, orc.apply_mpo, (L, M1, M2, R), num_restarts=2, krylov_dim=5,
                                                pro=True, pro_max_reorth=2)
        _, E_pro = parallel.lanczos_ground_state_sharded(shard_in, op, (L_rows, M1, M2, R), num_restarts=2,
                                                         krylov_dim=5, pro=True, pro_max_reorth=2,
                                                         _vector_ops=TorchVectorOps())
        # the returned shards assemble into a unit vector
        nrm2 = torch.tensor([float((v_sh.abs() ** 2).sum())], dtype=torch.float64)
        dist.all_reduce(nrm2)
        result_q.put((rank, err_mv, err_gather, E, E_ref, E_pro, E_pro_ref, float(nrm2)))
    except Exception as ex:      # fail fast in the parent instead of a queue timeout
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chi_l,chi_r,dtype_name", [(8, 6, "float64"), (7, 5, "complex128"), (16, 16, "complex128")])
def test_sharded_matvec_and_lanczos_world2(chi_l, chi_r, dtype_name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, chi_l, chi_r, dtype_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        item = q.get(timeout=240)
        if item[0] == "error":
            for p in procs:
                p.kill()
            pytest.fail(f"rank {item[1]} failed:\n{item[2]}")
        res.append(item)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Es = []
    for rank, err_mv, err_gather, E, E_ref, E_pro, E_pro_ref, nrm2 in res:
        assert err_mv <= 1e-13 and err_gather <= 1e-13
        assert abs(E - E_ref) <= 1e-10 * abs(E_ref)
        assert abs(E_pro - E_pro_ref) <= 1e-10 * abs(E_pro_ref)
        assert abs(nrm2 - 1.0) <= 1e-12
        Es.append(E)
    assert Es[0] == Es[1]                      # every rank sees the same Ritz value


def _dmrg_worker(rank, world, port, dtype_name, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import io
        import contextlib
        torch.set_num_threads(2)
        import quantum_simulation_b200 as qsb
        from quantum_simulation_b200 import parallel
        from oracle import dmrg_oracle as orc
        from conftest import golden_mpo
        from engine_cpu import CpuEngine
        dtype = getattr(torch, dtype_name)
        mpo = golden_mpo("heis_c128_L20" if dtype.is_complex else "heis_f64_L20", dtype)
        sched = dict(numsweeps=2, maxdim=[16, 16], krylov_dim=[4, 4], cutoff=[0.0, 1e-12], reortho=[False, True],
                     maxreortho=[1, 1])
        ref_mps = orc.OracleMPS(20, 2, 16, dtype, seed=5)
        E_ref, h_ref = orc.dmrg_run(ref_mps, mpo, **sched)
        with contextlib.redirect_stdout(io.StringIO()):
            mps = qsb.MPS(20, 2, 16, "cpu", dtype, seed=5)
        sw = qsb.Sweeps(2)
        sw.set_schedule(maxdim=sched["maxdim"], krylov_dim=sched["krylov_dim"], cutoff=sched["cutoff"],
                        reortho=sched["reortho"], maxreortho=sched["maxreortho"])
        dm = parallel.ShardedDMRG(mps, mpo, device="cpu", dtype=dtype, min_shard_chi=4, engine=CpuEngine())
        E, _ = dm(sw, dispon=0)
        rel = max(abs(a - b) / abs(b) for a, b in zip(E, E_ref))
        result_q.put((rank, len(E) == len(E_ref), rel, dm.history["bond_dims"] == h_ref["bond_dims"], dict(dm.comm),
                      float(E[-1])))
    except Exception:
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,world", [("float64", 2), ("complex128", 2), ("float64", 4)])
def test_sharded_dmrg_matches_unsharded(dtype_name, world):
    """Full two-site DMRG (2 sweeps, Heisenberg L=20, chi=16) through ShardedDMRG on 2 gloo ranks: row-sharded
    environments, sharded Lanczos, reduce-scatter env updates, broadcast split -- per-update energies equal the
    unsharded run to 1e-10 relative and every rank ends with the same energy."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dmrg_worker, args=(r, world, port, dtype_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        item = q.get(timeout=600)
        if item[0] == "error":
            for p in procs:
                p.kill()
            pytest.fail(f"rank {item[1]} failed:\n{item[2]}")
        res.append(item)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_len, rel, same_bonds, comm, E_last in res:
        assert same_len and same_bonds
        assert rel <= 1e-10, f"rank {rank}: max relative deviation {rel:.2e}"
        assert comm["sharded_steps"] > 0 and comm["replicated_steps"] > 0
        assert comm["reduce_scatter_env"] > 0 and comm["all_gather_env"] > 0 and comm["broadcast_split"] > 0
    assert len({r[5] for r in res}) == 1


def test_row_partition():
    from quantum_simulation_b200.parallel import row_partition
    assert row_partition(8, 2) == [(0, 4), (4, 8)]
    assert row_partition(7, 2) == [(0, 4), (4, 7)]
    assert row_partition(3, 8) == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert row_partition(2048, 8)[7] == (1792, 2048)
    with pytest.raises(ValueError):
        row_partition(4, 0)


def test_fused_exchange_eligibility():
    """the fused (arrival-gated, TMA-staged) matvec is only chosen for bonds the TMA can describe;
    truncated / spin-1 right bonds fall back to all-gather + plain matvec instead of aborting the sweep."""
    import torch
    from quantum_simulation_b200.parallel import fused_eligible
    f64, c128 = torch.float64, torch.complex128
    assert fused_eligible(2048, 4 * 2048, f64, 8) and fused_eligible(512, 16 * 512, c128, 8)
    assert not fused_eligible(64, 4 * 50, f64, 2)          # chi_r = 50: 200 is not a multiple of 16
    assert not fused_eligible(64, 9 * 27, f64, 2)          # spin-1, chi_r = 27
    assert fused_eligible(64, 4 * 50, c128, 2)             # 200 % 8 == 0 for complex128
    assert not fused_eligible(72, 4 * 64, f64, 8)          # k range does not split into whole k-tiles per rank
    assert not fused_eligible(64, 4 * 64, c128, 16) and fused_eligible(128, 4 * 64, c128, 16)


# ---------------------------------------------------------------------------------------------------------------
# two-site split on row shards, noise, sharded theta build (f.1 / f.2 / missing #5)
# ---------------------------------------------------------------------------------------------------------------
def _spawn(target, world, args, timeout=600):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + tuple(args) + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        item = q.get(timeout=timeout)
        if item[0] == "error":
            for p in procs:
                p.kill()
            pytest.fail(f"rank {item[1]} failed:\n{item[2]}")
        res.append(item)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda x: x[0])


def _split_worker(rank, world, port, dtype_name, m, n, k, decay, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from quantum_simulation_b200 import svd as qsvd
        dtype = getattr(torch, dtype_name)
        g = torch.Generator().manual_seed(3)
        r = min(m, n)
        U0, _ = torch.linalg.qr(torch.randn(m, r, dtype=dtype, generator=g))
        V0, _ = torch.linalg.qr(torch.randn(n, r, dtype=dtype, generator=g))
        sv = torch.exp(-torch.arange(r, dtype=torch.float64) * decay)
        theta = (U0 * sv.to(dtype)) @ V0.conj().T
        theta = theta / torch.linalg.norm(theta)
        assert qsvd.sharded_split_eligible(m, n, k, world) == (decay is not None)
        mr = m // world
        out = qsvd.sketch_split_sharded(theta[rank * mr:(rank + 1) * mr].contiguous(), m, k, 0.0, dist.group.WORLD, {})
        if out is None:
            again = qsvd.sketch_split_sharded(theta[rank * mr:(rank + 1) * mr].contiguous(), m, k, 0.0, dist.group.WORLD)
            result_q.put((rank, "rejected", qsvd.stats["last_rho"], again is None, dict(qsvd.stats)))
            return
        U, S, Vh, keep = out
        Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
        best = (Ue[:, :k] * Se[:k].to(dtype)) @ Vhe[:k]
        eye = torch.eye(keep, dtype=dtype)
        result_q.put((rank, "ok", keep,
                      float(torch.linalg.norm(U.conj().T @ U - eye)), float(torch.linalg.norm(Vh @ Vh.conj().T - eye)),
                      float((S[:keep] ** 2 - Se[:keep] ** 2).abs().max()),
                      float(torch.linalg.norm((U * S[:keep].to(dtype)) @ Vh - best)),
                      abs(float((S[keep:] ** 2).sum()) - float((Se[keep:] ** 2).sum())),
                      U.numpy().tobytes() + S.numpy().tobytes() + Vh.numpy().tobytes(), dict(qsvd.stats)))
    except Exception:
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,world,shape", [("float64", 2, (256, 256)), ("complex128", 2, (256, 320)),
                                                    ("float64", 4, (320, 256))])
def test_two_site_split_on_row_shards(dtype_name, world, shape):
    """svd.sketch_split_sharded on gloo ranks: nobody holds the full block, yet every rank ends with the SAME factors
    (bitwise), isometric to 1e-12, squared Schmidt values and discarded weight within 2e-13 of the SVD's, and the
    truncated block equal to the optimally truncated one to 1e-7 (Schmidt values below 1e-8 are noise in any Gram
    method)."""
    res = _spawn(_split_worker, world, (dtype_name, shape[0], shape[1], 64, 0.3))
    for rank, status, keep, du, dv, ds, drec, dS, blob, stats in res:
        assert status == "ok" and keep == 64
        assert du < 1e-12 and dv < 1e-12 and ds <= 2e-13 and drec < 1e-7 and dS <= 2e-13
        assert stats.get("sketch_refined", 0) == 0
    assert len({r[8] for r in res}) == 1


def test_two_site_split_on_row_shards_refines_a_marginal_miss():
    """A spectrum whose plain compression misses the bound by a small factor: all ranks take the subspace-iteration
    refinement together (V = orth(B^H) on column shards, Y <- theta V) and then accept; same guarantees as above."""
    res = _spawn(_split_worker, 2, ("float64", 256, 256, 64, 0.2))
    for rank, status, keep, du, dv, ds, drec, dS, blob, stats in res:
        assert status == "ok" and keep == 64
        assert du < 1e-12 and dv < 1e-12 and ds <= 2e-13 and drec < 1e-6 and dS <= 2e-13
        assert stats["sketch_refined"] == 1 and stats["last_rho"] <= 1e-13
    assert len({r[8] for r in res}) == 1


def test_two_site_split_on_row_shards_rejects_collectively():
    """Slow decay: all ranks reject together (one all-reduced rho), and all skip the next attempt together."""
    res = _spawn(_split_worker, 2, ("float64", 256, 256, 64, 0.02))
    for rank, status, rho, again_none, stats in res:
        assert status == "rejected" and rho > 1e-13 and again_none
        assert stats["sketch_rejected"] == 1 and stats["sketch_skipped"] == 1


def _dmrg_split_noise_worker(rank, world, port, dtype_name, noise, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import io
        import contextlib
        torch.set_num_threads(2)
        import quantum_simulation_b200 as qsb
        from quantum_simulation_b200 import parallel, svd as qsvd
        from conftest import golden_mpo
        from engine_cpu import CpuEngine
        # make the chi = 16 bulk blocks (32 x 32, keep 16) eligible for the sharded split: the compression is then the
        # full range (l = 32), i.e. exact, which is what lets the energies be compared at 1e-10
        qsvd.FAST_MIN, qsvd.SKETCH_MAX_FRACTION = 8, 1.01
        dtype = getattr(torch, dtype_name)
        mpo = golden_mpo("heis_c128_L20" if dtype.is_complex else "heis_f64_L20", dtype)
        with contextlib.redirect_stdout(io.StringIO()):
            mps = qsb.MPS(20, 2, 16, "cpu", dtype, seed=5)
        sw = qsb.Sweeps(2)
        sw.set_schedule(maxdim=[16, 16], krylov_dim=[4, 4], cutoff=[0.0, 0.0], noise=[noise, 0.0])
        dm = parallel.ShardedDMRG(mps, mpo, device="cpu", dtype=dtype, min_shard_chi=4, engine=CpuEngine())
        torch.manual_seed(11)                  # rank 0's global RNG is the one the noise comes from
        E, _ = dm(sw, dispon=0)
        result_q.put((rank, [float(e) for e in E], dict(dm.comm), list(dm.history["bond_dims"])))
    except Exception:
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,noise", [("float64", 0.0), ("complex128", 0.0), ("float64", 1e-3)])
def test_sharded_dmrg_split_on_shards_and_noise(dtype_name, noise):
    """ShardedDMRG on 2 ranks with the two-site split running on the row shards (and, in the last case, the reference's
    noise term drawn on rank 0) against the same driver on ONE rank, where every step is the replicated reference
    formula: per-update energies within 1e-10 relative, same bond dimensions, sharded splits actually used."""
    one = _spawn(_dmrg_split_noise_worker, 1, (dtype_name, noise))[0]
    two = _spawn(_dmrg_split_noise_worker, 2, (dtype_name, noise))
    E1 = np.array(one[1])
    for rank, E, comm, bonds in two:
        E = np.array(E)
        assert len(E) == len(E1) and bonds == one[3]
        assert np.max(np.abs(E - E1) / np.abs(E1)) <= 1e-10
        assert comm.get("sharded_splits", 0) > 0 and comm["all_reduce_split"] > 0
        assert comm.get("sharded_gram_splits", 0) > 0            # non-truncating / growing blocks: exact split on shards
        assert (comm.get("broadcast_noise", 0) > 0) == (noise > 0)
    assert two[0][1] == two[1][1]


def _dmrg_shard_mps_worker(rank, world, port, dtype_name, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import io
        import contextlib
        torch.set_num_threads(2)
        import quantum_simulation_b200 as qsb
        from quantum_simulation_b200 import parallel
        from conftest import golden_mpo
        from engine_cpu import CpuEngine
        dtype = getattr(torch, dtype_name)
        mpo = golden_mpo("heis_c128_L20" if dtype.is_complex else "heis_f64_L20", dtype)
        out = {}
        for shard in (False, True):
            with contextlib.redirect_stdout(io.StringIO()):
                mps = qsb.MPS(20, 2, 16, "cpu", dtype, seed=5)
            sw = qsb.Sweeps(2)
            sw.set_schedule(maxdim=[16, 16], krylov_dim=[4, 4], cutoff=[0.0, 1e-12])
            dm = parallel.ShardedDMRG(mps, mpo, device="cpu", dtype=dtype, min_shard_chi=4, engine=CpuEngine(),
                                      shard_mps=shard)
            E, _ = dm(sw, dispon=0)
            stored = sum(t.numel() for t in mps.A) + sum(t.numel() for t in mps.B)
            n_sharded = len(dm._A.left_dim) + len(dm._B.left_dim)
            if shard:                      # a second schedule on the still-sharded MPS, then reassemble
                dm2 = parallel.ShardedDMRG(mps, mpo, device="cpu", dtype=dtype, min_shard_chi=4, engine=CpuEngine(),
                                           shard_mps=True)
                sw1 = qsb.Sweeps(1)
                sw1.set_schedule(maxdim=[16], krylov_dim=[4], cutoff=[1e-12])
                E2, _ = dm2(sw1, dispon=0)
                dm2.gather_mps()
                out["E2_shard"] = [float(e) for e in E2]
                out["gathered"] = sum(t.numel() for t in mps.A) + sum(t.numel() for t in mps.B)
                out["norm_center"] = float(torch.linalg.norm(mps.B[0]))
            else:
                sw1 = qsb.Sweeps(1)
                sw1.set_schedule(maxdim=[16], krylov_dim=[4], cutoff=[1e-12])
                E2, _ = parallel.ShardedDMRG(mps, mpo, device="cpu", dtype=dtype, min_shard_chi=4,
                                             engine=CpuEngine())(sw1, dispon=0)
                out["E2_full"] = [float(e) for e in E2]
            out[shard] = ([float(e) for e in E], stored, n_sharded, list(dm.history["bond_dims"]), dict(dm.comm))
        result_q.put((rank, out))
    except Exception:
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float64", "complex128"])
def test_sharded_dmrg_with_row_sharded_mps(dtype_name):
    """ShardedDMRG(shard_mps=True): the site tensors themselves are kept as row blocks (what makes chi = 4096, d = 4
    fit: 137 GB of replicated tensors -> 17 GB per GPU) and the one full tensor a bond step needs is all-gathered on
    demand.  Same energies (bitwise) and bond dimensions as with the replicated MPS, about half the stored elements on
    2 ranks, a second schedule runs on the still-sharded state, gather_mps() restores the reference's layout."""
    res = _spawn(_dmrg_shard_mps_worker, 2, (dtype_name,))
    for rank, out in res:
        E_full, stored_full, n0, bonds_full, _ = out[False]
        E_sh, stored_sh, n1, bonds_sh, comm = out[True]
        assert E_sh == E_full and bonds_sh == bonds_full
        assert n0 == 0 and n1 > 20 and stored_sh < 0.62 * stored_full
        assert comm["all_gather_site"] > 0
        assert out["E2_shard"] == out["E2_full"]
        assert out["gathered"] >= stored_full * 0.9 and abs(out["norm_center"] - 1.0) < 1e-10
    assert res[0][1][True][0] == res[1][1][True][0]


def _gram_split_worker(rank, world, port, dtype_name, m, n, k, cutoff, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from quantum_simulation_b200 import svd as qsvd
        dtype = getattr(torch, dtype_name)
        g = torch.Generator().manual_seed(13)
        r = min(m, n)
        U0, _ = torch.linalg.qr(torch.randn(m, r, dtype=dtype, generator=g))
        V0, _ = torch.linalg.qr(torch.randn(n, r, dtype=dtype, generator=g))
        sv = torch.exp(-torch.arange(r, dtype=torch.float64) * 0.05)          # slow decay: real truncation
        theta = (U0 * sv.to(dtype)) @ V0.conj().T
        theta = theta / torch.linalg.norm(theta)
        mr = m // world
        U, S, Vh, keep = qsvd.gram_split_sharded(theta[rank * mr:(rank + 1) * mr].contiguous(), m, k, cutoff,
                                                 dist.group.WORLD, {})
        Ue, Se, Vhe = torch.linalg.svd(theta, full_matrices=False)
        kr = qsvd._keep_from_spectrum(Se, k, cutoff)
        eye = torch.eye(keep, dtype=dtype)
        best = (Ue[:, :kr] * Se[:kr].to(dtype)) @ Vhe[:kr]
        result_q.put((rank, keep, kr, float(torch.linalg.norm(U.conj().T @ U - eye)),
                      float(torch.linalg.norm(Vh @ Vh.conj().T - eye)), float((S[:keep] - Se[:keep]).abs().max()),
                      float(torch.linalg.norm((U * S[:keep].to(dtype)) @ Vh - best)),
                      abs(float((S[keep:] ** 2).sum()) - float((Se[keep:] ** 2).sum())),
                      U.numpy().tobytes() + S.numpy().tobytes() + Vh.numpy().tobytes()))
    except Exception:
        import traceback
        result_q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,world,shape,k,cutoff", [("float64", 2, (96, 96), 40, 0.0), ("complex128", 2, (128, 64), 30, 0.0),
                                                             ("float64", 4, (128, 96), 10 ** 6, 1e-8)])
def test_exact_gram_split_on_row_shards(dtype_name, world, shape, k, cutoff):
    """svd.gram_split_sharded (the exact path for blocks with real truncation, theta never gathered): the reference's
    keep, isometric factors, the optimally truncated block to 1e-7, identical factors on every rank."""
    res = _spawn(_gram_split_worker, world, (dtype_name, shape[0], shape[1], k, cutoff))
    for rank, keep, kr, du, dv, ds, drec, dw, blob in res:
        assert keep == kr
        assert du < 1e-12 and dv < 1e-12 and ds < 1e-9 and drec < 1e-7 and dw < 1e-13
    assert len({r[8] for r in res}) == 1

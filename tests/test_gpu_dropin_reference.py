"""The drop-in claim, exercised through the REFERENCE'S OWN classes (missing #3): the unmodified
reference staged under baseline/_ref/ (baseline/stage_reference.py) has its three contraction modules routed to the
C ABI by the binding of INTEGRATION.md section B (integration/qsb200_binding.py -- plain ctypes, not this repo's
Python package), and then ITS ``DMRG`` driver, ``EnvironmentManager``, ``Sweeps``, ``MPS`` and
``lanczos_ground_state`` run on ``device='cuda'`` against the golden energies of the stock reference."""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, golden_mpo, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref():
    stage = _load(os.path.join(ROOT, "baseline", "stage_reference.py"), "_stage_reference")
    if not stage.staged():
        pytest.skip("reference not staged under baseline/_ref (run baseline/stage_reference.py in the build container)")
    qs = stage.import_reference()
    assert "/root/reference" not in os.path.abspath(qs.__file__)
    binding = _load(os.path.join(ROOT, "integration", "qsb200_binding.py"), "_qsb200_binding")
    import quantum_simulation.dmrg as rdmrg
    undo = binding.install(rdmrg)
    yield qs, rdmrg, binding
    undo()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _launches():
    import quantum_simulation_b200 as qsb          # only for the library's launch counters
    return qsb.launch_counters()["total"]


@pytest.mark.parametrize("name", ["nb02_tfim_L6", "nb03_heis_L4", "cal_heis_L20", "cal_tfim_L40",
                                  "c1_xxz_L20_chi64_c128", "c1_xxz_L20_chi64_f64", "heis_f64_L20_pro"])
def test_reference_dmrg_runs_on_the_kernels(ref, name):
    qs, rdmrg, binding = ref
    from quantum_simulation.mps import MPS
    z = load_golden("dmrg")
    L, d, chi0, seed, ns = (int(x) for x in z[f"{name}/cfg"])
    dt = torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64
    mps = quiet(MPS, L, d, chi0, "cpu", dt, seed).to(DEV)
    mpo = golden_mpo(str(z[f"{name}/mpo"]), dt, DEV)
    sw = rdmrg.Sweeps(ns)
    sw.set_schedule(maxdim=[int(x) for x in z[f"{name}/maxdim"]], krylov_dim=[int(x) for x in z[f"{name}/krylov_dim"]],
                    cutoff=[float(x) for x in z[f"{name}/cutoff"]], reortho=[bool(x) for x in z[f"{name}/reortho"]],
                    maxreortho=[int(x) for x in z[f"{name}/maxreortho"]], noise=[0.0] * ns)
    before = _launches()
    dm = quiet(rdmrg.DMRG, mps, mpo, device=DEV, dtype=dt, contract_backend="ncon")      # the reference's driver
    E, _ = quiet(dm, sw, dispon=0, maxit=2)
    assert _launches() - before > 10 * len(E), "the reference's modules did not reach the C ABI"
    E, Eg = np.array(E), z[f"{name}/energies"]
    assert len(E) == len(Eg)
    assert dm.history["bond_dims"] == [int(x) for x in z[f"{name}/bond_dims"]]
    rel = np.max(np.abs(E - Eg) / np.abs(Eg))
    assert rel < 1e-10, f"max relative energy deviation {rel:.2e}"


def test_reference_lanczos_calls_the_binding_with_out(ref):
    """lanczos_solver.py:234-238: ``linear_operator(v, *args, out=w)`` -- the patched module writes into the solver's
    workspace buffer and returns the flat view, and the Ritz value matches the reference on the CPU."""
    qs, rdmrg, binding = ref
    from quantum_simulation.eigensolver import lanczos_ground_state
    dt = torch.complex128
    Ms = golden_mpo("heis_c128_L20", dt)
    g = torch.Generator().manual_seed(5)
    chi = 24
    L = torch.randn(5, chi, chi, dtype=dt, generator=g)
    R = torch.randn(5, chi, chi, dtype=dt, generator=g)
    L, R = L + L.conj().transpose(1, 2), R + R.conj().transpose(1, 2)
    v0 = torch.randn(chi * 4 * chi, dtype=dt, generator=g)
    op = rdmrg.ApplyMPO("ncon")
    _, E_cpu = lanczos_ground_state(v0, op, (L, Ms[9], Ms[10], R), num_restarts=2, krylov_dim=4)
    before = _launches()
    args = tuple(t.to(DEV) for t in (L, Ms[9], Ms[10], R))
    v, E_gpu = lanczos_ground_state(v0.to(DEV), op, args, num_restarts=2, krylov_dim=4)
    assert _launches() - before == 8 * 4                       # 8 matvecs x (prepare, GEMM, mix, GEMM)
    assert abs(E_gpu - E_cpu) <= 1e-10 * abs(E_cpu)
    w = torch.empty_like(v)
    res = op(v, *args, out=w)
    assert res.data_ptr() == w.data_ptr()
    assert rel_err(w, op(v.cpu(), L, Ms[9], Ms[10], R)) <= 1e-12
    with pytest.raises(ValueError, match="out shape/device/dtype mismatch"):
        op(v, *args, out=torch.empty(3, dtype=dt, device=DEV))


def test_reference_modules_keep_their_cpu_path(ref):
    """Non-CUDA tensors still take the reference's own code (the binding adds a branch, it does not remove one)."""
    qs, rdmrg, binding = ref
    dt = torch.float64
    M = golden_mpo("heis_f64_L20", dt)[9]
    g = torch.Generator().manual_seed(9)
    Lp = torch.randn(5, 6, 6, dtype=dt, generator=g)
    A = torch.randn(6, 2, 6, dtype=dt, generator=g)
    before = _launches()
    out_cpu = rdmrg.LeftEnvUpdate("ncon")(Lp, M, A)
    assert _launches() == before and out_cpu.device.type == "cpu"
    out_gpu = rdmrg.LeftEnvUpdate("ncon")(Lp.to(DEV), M.to(DEV), A.to(DEV))
    assert _launches() > before and rel_err(out_gpu, out_cpu) <= 1e-12

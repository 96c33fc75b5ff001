"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/qsb200.h declares; the
host-side mirror has the reference's API surface and error behaviour; nothing computes without a GPU."""
import os
import re

import pytest
import torch

from conftest import ROOT
import quantum_simulation_b200 as qsb
from quantum_simulation_b200 import _lib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qsb200.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:int|double|size_t|void|const char\*)\s+(qsb_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert len(declared) >= 19
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in qsb200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.qsb_version() >= 200
    import ctypes as C
    for fn, st in ((lib.qsb_sizeof_heff_desc, _lib.HeffDesc), (lib.qsb_sizeof_env_desc, _lib.EnvDesc),
                   (lib.qsb_sizeof_gemm_desc, _lib.GemmDesc)):
        assert fn() == C.sizeof(st)          # the ctypes mirrors are byte-identical to include/qsb200.h
    assert _lib.status_string(0) == "ok" and "workspace" in _lib.status_string(5)
    c = qsb.launch_counters(reset=True)
    assert c["total"] == 0


def test_descriptor_validation_without_gpu():
    lib = _lib.load()
    import ctypes as C
    d = _lib.HeffDesc()
    d.dtype = 0
    nb = C.c_size_t(0)
    assert lib.qsb_heff_workspace_bytes(C.byref(d), C.byref(nb)) == 1          # zero extents -> QSB_ERR_SHAPE
    for f in ("chi_l_out", "chi_l_in", "chi_r_out", "chi_r_in"):
        setattr(d, f, 64)
    for f in ("d1_out", "d1_in", "d2_out", "d2_in"):
        setattr(d, f, 2)
    d.W_l = d.W_m = d.W_r = 5
    d.L_stride = (C.c_int64 * 3)(64 * 64, 64, 1)
    d.R_stride = (C.c_int64 * 3)(64 * 64, 64, 1)
    assert lib.qsb_heff_workspace_bytes(C.byref(d), C.byref(nb)) == 0
    assert nb.value >= 2 * 5 * 4 * 64 * 64 * 8
    flops = lib.qsb_heff_flops(C.byref(d))
    assert flops == 2.0 * (2 * 5 * 64 ** 3 * 4 + 2 * 10 * 10 * 2 * 64 * 64)
    d.dtype = 7
    assert lib.qsb_heff_workspace_bytes(C.byref(d), C.byref(nb)) == 2          # QSB_ERR_DTYPE
    assert lib.qsb_heff_apply(C.byref(d), None) == 2
    e = _lib.EnvDesc()
    assert lib.qsb_env_workspace_bytes(C.byref(e), C.byref(nb)) == 1


def test_api_surface_matches_reference():
    for name in ("ApplyMPO", "LeftEnvUpdate", "RightEnvUpdate", "EnvironmentManager", "Sweeps", "DMRG", "MPS",
                 "lanczos_ground_state", "set_contract_backend", "random_mps"):
        assert hasattr(qsb, name)
    import inspect
    sig = inspect.signature(qsb.lanczos_ground_state)
    assert list(sig.parameters) == ["initial_vector", "linear_operator", "operator_args", "num_restarts",
                                    "krylov_dim", "pro", "pro_max_reorth", "beta_tol", "ritz_residual_tol"]
    assert sig.parameters["num_restarts"].default == 2 and sig.parameters["krylov_dim"].default == 4
    assert sig.parameters["beta_tol"].default == 1e-12
    assert list(inspect.signature(qsb.ApplyMPO.forward).parameters) == ["self", "psi_flat", "L", "M1", "M2", "R", "out"]
    assert list(inspect.signature(qsb.LeftEnvUpdate.forward).parameters) == ["self", "Lp", "M", "Ap"]
    assert list(inspect.signature(qsb.RightEnvUpdate.forward).parameters) == ["self", "M", "Rnext", "Bp1"]


def test_sweeps_schedule_errors():
    with pytest.raises(ValueError):
        qsb.Sweeps(0)
    s = qsb.Sweeps(3)
    assert s.maxdim == [10, 10, 10] and s.krylov_dim == [4, 4, 4] and s.cutoff == [1e-12] * 3
    s.set_schedule(maxdim=[4, 8, 12], reortho=[False, False, True])
    assert s.maxdim == [4, 8, 12]
    with pytest.raises(AttributeError):
        s.set_schedule(nope=[1, 2, 3])
    with pytest.raises(ValueError):
        s.set_schedule(maxdim=[1, 2])
    assert "maxdim: [4, 8, 12]" in repr(s)


def test_backend_names():
    for b in ("auto", "einsum", "ncon"):
        qsb.ApplyMPO(b)
    with pytest.raises(ValueError):
        qsb.ApplyMPO("cutensor")
    with pytest.raises(ValueError):
        qsb.set_contract_backend("nope", "cuda")


def test_no_cpu_fallback():
    L = torch.ones(1, 1, 1, dtype=torch.float64)
    M = torch.ones(1, 1, 2, 2, dtype=torch.float64)
    with pytest.raises(RuntimeError):          # NotImplementedError is a RuntimeError; never TypeError
        qsb.ApplyMPO()(torch.ones(4, dtype=torch.float64), L, M, M, L)      # operator on the CPU: no kernel
    with pytest.raises(RuntimeError):
        qsb.lanczos_ground_state(torch.ones(4, dtype=torch.float64), lambda x, A: A @ x, (torch.eye(4, dtype=torch.float64),))
    with pytest.raises(RuntimeError):
        mps = qsb.MPS(4, 2, 4, "cpu", torch.float64, seed=1)
        qsb.DMRG(mps, [M] * 4, device="cpu")


def test_environment_manager_boundaries_and_errors():
    mpo = [torch.zeros(1, 3, 2, 2), torch.zeros(3, 3, 2, 2), torch.zeros(3, 1, 2, 2)]
    env = qsb.EnvironmentManager(mpo, "cpu", torch.float64)
    assert tuple(env.get_L(0).shape) == (1, 1, 1) and tuple(env.get_R(3).shape) == (1, 1, 1)
    with pytest.raises(RuntimeError, match=r"L\[1\] requested"):
        env.get_L(1)
    with pytest.raises(RuntimeError, match=r"R\[0\] requested"):
        env.get_R(0)


def test_mps_matches_reference_init():
    """Our MPS built on the CPU with a seed is the reference's (golden init_abs_sums from the real MPS class)."""
    import numpy as np
    from conftest import load_golden
    z = load_golden("dmrg")
    for name in ("nb02_tfim_L6", "cal_heis_L20", "c1_xxz_L20_chi64_f64"):
        L, d, chi0, seed, _ = (int(x) for x in z[f"{name}/cfg"])
        dt = torch.complex128 if str(z[f"{name}/dtype"]) == "c128" else torch.float64
        mps = qsb.MPS(L, d, chi0, "cpu", dt, seed=seed)
        sums = np.array([t.detach().abs().sum().item() for t in mps.B])
        assert np.allclose(sums, z[f"{name}/init_abs_sums"], rtol=1e-13, atol=0)
        assert mps.ortho_center == 0 and len(mps.sWeight) == L + 1


def test_mps_position_keeps_state_and_canonical_form():
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        mps = qsb.MPS(7, 2, 6, "cpu", torch.complex128, seed=9)

        def dense(m):
            v = m.A[0] if m.ortho_center == 0 else None
            ts = [m.A[p] if p <= m.ortho_center else m.B[p] for p in range(m.Nsites)]
            out = ts[0]
            for t in ts[1:]:
                out = torch.tensordot(out, t, dims=([out.dim() - 1], [0]))
            return out.reshape(-1)
        psi0 = dense(mps)
        for target in (4, 1, 6, 0):
            mps.position(target)
            assert mps.ortho_center == target
            psi = dense(mps)
            assert abs(abs(torch.vdot(psi0, psi)) - 1.0) < 1e-12          # same state up to a phase
            for p in range(target):                                      # left of the centre: left isometries
                a = mps.A[p].reshape(-1, mps.A[p].shape[2])
                assert torch.linalg.norm(a.conj().T @ a - torch.eye(a.shape[1], dtype=a.dtype)) < 1e-12
            for p in range(target + 1, mps.Nsites):                      # right of it: right isometries
                b = mps.B[p].reshape(mps.B[p].shape[0], -1)
                assert torch.linalg.norm(b @ b.conj().T - torch.eye(b.shape[0], dtype=b.dtype)) < 1e-12
    with pytest.raises(ValueError):
        mps.position(7)


def test_c_abi_from_plain_c(tmp_path):
    """Compile and run a plain-C consumer of include/qsb200.h against the built library (gcc, no GPU needed)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(_lib.LIB_PATH)
    _lib.load()
    exe = str(tmp_path / "abi_check")
    cmd = [gcc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-L", libdir, "-lqsb200", f"-Wl,-rpath,{libdir}", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi_check ok" in run.stdout


def test_dtype_mismatch_is_a_value_error():
    """Mixed dtypes must be rejected before any pointer reaches the kernels (CPU tensors suffice to check)."""
    L = torch.ones(1, 1, 1, dtype=torch.complex128)
    M = torch.ones(1, 1, 2, 2, dtype=torch.float64)
    with pytest.raises(ValueError, match="dtype"):
        qsb.ApplyMPO()(torch.ones(4, dtype=torch.float64), L, M, M, L.real)
    with pytest.raises(ValueError, match="dtype"):
        qsb.LeftEnvUpdate()(L, M, torch.ones(1, 2, 1, dtype=torch.complex128))
    with pytest.raises(ValueError, match="dtype"):
        qsb.RightEnvUpdate()(M, L, torch.ones(1, 2, 1, dtype=torch.complex128))
    with pytest.raises(ValueError, match="dtype"):
        qsb.ApplyMPO()(torch.ones(4, dtype=torch.float32), L.real.float(), M.float(), M.float(), L.real.float())


def test_mps_constructor_and_position_match_reference_fixture():
    """tests/golden/mps.npz holds the unmodified reference's MPS tensors after construction and after a sequence of
    position() moves; the mirror reproduces them bit for bit on the CPU (same torch ops in the same order)."""
    import io
    import contextlib
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "mps.npz"))
    for tag, dt in (("c128", torch.complex128), ("f64", torch.float64)):
        N, d, chi, seed = (int(x) for x in z[f"{tag}/cfg"])
        with contextlib.redirect_stdout(io.StringIO()):
            m = qsb.MPS(N, d, chi, "cpu", dt, seed=seed)
            for p in range(N):
                assert np.array_equal(m.B[p].detach().numpy(), z[f"{tag}/init_B{p}"])
            for tgt in z[f"{tag}/moves"]:
                m.position(int(tgt))
        for p in range(N):
            assert np.allclose(m.A[p].detach().numpy(), z[f"{tag}/A{p}"], rtol=0, atol=1e-15)
            assert np.allclose(m.B[p].detach().numpy(), z[f"{tag}/B{p}"], rtol=0, atol=1e-15)
        for p in range(N + 1):
            assert np.allclose(m.sWeight[p].detach().numpy(), z[f"{tag}/s{p}"], rtol=0, atol=1e-15)


def test_documented_binding_matches_the_header():
    """INTEGRATION.md's stub stopped short of the real struct.  Now: the standalone binding
    (integration/qsb200_binding.py), the struct text quoted in INTEGRATION.md, this package's own mirror and the
    compiled library all agree, field for field and byte for byte."""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location("_qsb200_binding", os.path.join(ROOT, "integration", "qsb200_binding.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib = b.load()                                            # asserts sizeof against qsb_sizeof_*()
    for mine, theirs in ((_lib.HeffDesc, b.HeffDesc), (_lib.EnvDesc, b.EnvDesc)):
        assert [(n, C.sizeof(t)) for n, t in mine._fields_] == [(n, C.sizeof(t)) for n, t in theirs._fields_]
        assert C.sizeof(mine) == C.sizeof(theirs)
    assert lib.qsb_sizeof_heff_desc() == C.sizeof(b.HeffDesc)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    hdr = open(os.path.join(ROOT, "include", "qsb200.h")).read()
    for st, cname in ((b.HeffDesc, "qsb_heff_desc"), (b.EnvDesc, "qsb_env_desc")):
        names = [n for n, _ in st._fields_]
        doc_names = re.findall(r'\("(\w+)", C\.', doc[doc.index(f"== {cname}"):].split("]\n\n")[0])
        assert doc_names == names, f"INTEGRATION.md stub of {cname} drifted"
        body = hdr[:hdr.index(f"}} {cname};")]
        body = body[body.rindex("typedef struct {"):]
        at = 0                                                # every field appears in the header, in this order
        for n in names:
            m = re.compile(rf"\b{n}\b").search(body, at)
            assert m is not None, f"{cname}.{n} missing from (or out of order in) include/qsb200.h"
            at = m.end()


def test_mps_long_chain_overflow_is_the_references_and_is_flagged():
    """The reference's constructor never rescales the carry of its QR sweep (mps.py:52-60): on long chains the norm of
    site 0 overflows and the centre tensor comes out zero / non-finite (L=100, chi=2048 in BASELINE config 3 -- the cause
    of the run-to-run drift of the r01 sweep, profiles/r02_determinism.md).  The mirror reproduces that, warns, and offers
    rescale_carry=True (same ray, finite centre)."""
    import contextlib
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.warns(RuntimeWarning, match="global RNG"):
            bad = qsb.MPS(300, 2, 64, "cpu", torch.float64, 1)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            good = qsb.MPS(300, 2, 64, "cpu", torch.float64, 1, rescale_carry=True)
            short = qsb.MPS(20, 2, 16, "cpu", torch.float64, 1)                 # no overflow, no warning
    assert not bool(torch.isfinite(bad.B[0]).all()) or not bool(bad.B[0].any())
    assert abs(float(torch.linalg.norm(good.B[0])) - 1.0) < 1e-12
    for p in (1, 150, 299):                                                      # right-canonical away from the centre
        b = good.B[p].reshape(good.B[p].shape[0], -1)
        assert float(torch.linalg.norm(b @ b.T - torch.eye(b.shape[0], dtype=b.dtype))) < 1e-12
    assert abs(float(torch.linalg.norm(short.B[0])) - 1.0) < 1e-12


def test_host_side_observable_helpers():
    """ground_state_energy / energy_density / truncation_errors / excitation_gap of the reference's observables module
    (dmrg_observables.py:27-117, 209-233): no device work, same values and exception types."""
    assert qsb.ground_state_energy([1.0, -2.5 + 0j]) == -2.5
    assert qsb.ground_state_energy(torch.tensor([[3.0, 4.0]])) == 4.0
    with pytest.raises(ValueError):
        qsb.ground_state_energy([])
    assert qsb.energy_density(torch.tensor(-8.0), 20) == -0.4
    with pytest.raises(ValueError):
        qsb.energy_density(1.0, 0)
    assert qsb.truncation_errors({"discarded_weights": [1e-9, torch.tensor(2e-9, dtype=torch.float64)]}) == [1e-9, 2e-9]
    with pytest.raises(KeyError):
        qsb.truncation_errors({})
    with pytest.raises(TypeError):
        qsb.truncation_errors([1.0])
    with pytest.raises(NotImplementedError):
        qsb.excitation_gap()


def test_row_block_bounds_of_the_host_pipelines():
    """Result row blocks of the host-resident matvecs (ApplyMPO._forward_host, FusedShardedApplyMPO.forward_host): a
    partition of the rows, halving sizes, never below one GEMM tile row, at most 5 blocks."""
    from quantum_simulation_b200 import _lib
    assert _lib.row_block_bounds(2048) == [0, 1024, 1536, 1792, 1920, 2048]
    assert _lib.row_block_bounds(1024) == [0, 512, 768, 896, 1024]
    assert _lib.row_block_bounds(256) == [0, 128, 256]
    assert _lib.row_block_bounds(128) == [0, 128] and _lib.row_block_bounds(130) == [0, 130]
    assert _lib.row_block_bounds(1) == [0, 1] and _lib.row_block_bounds(96) == [0, 96]
    for rows in (64, 200, 384, 512, 640, 4096, 6144):
        b = _lib.row_block_bounds(rows)
        sizes = [y - x for x, y in zip(b[:-1], b[1:])]
        assert b[0] == 0 and b[-1] == rows and sum(sizes) == rows and len(sizes) <= 5
        assert all(s >= 128 for s in sizes) or len(sizes) == 1
        assert sizes == sorted(sizes, reverse=True)

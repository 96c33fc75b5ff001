"""GPU parity of the Lanczos solver: golden energies of the unmodified reference, and the semantics the
reference's own tests pin (eigensolver/test_lanczos.py:81-340) re-stated against our lanczos_ground_state.
Tolerance: 1e-10 absolute on 64-bit energies (test_lanczos.py:27-41)."""
import pytest
import torch

from conftest import load_golden, rel_err
import quantum_simulation_b200 as qsb
from quantum_simulation_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def matvec(x, A):
    return A @ x


def rand_herm(n, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, dtype=dtype, generator=g)
    return 0.5 * (A + A.conj().T) + 0.1 * torch.eye(n, dtype=dtype)


def atol_for(dtype):
    return 1e-10 if dtype in (torch.float64, torch.complex128) else 2e-4


@pytest.mark.parametrize("name", [str(s) for s in load_golden("lanczos")["names"]])
def test_lanczos_golden(name):
    z = load_golden("lanczos")
    H, v0, v = (torch.from_numpy(z[f"{name}/{k}"]) for k in ("H", "v0", "v"))
    k, r, pro, mr = (int(x) for x in z[f"{name}/params"])
    vec, E = qsb.lanczos_ground_state(v0.to(DEV), matvec, (H.to(DEV),), num_restarts=r, krylov_dim=k,
                                      pro=bool(pro), pro_max_reorth=mr)
    assert isinstance(E, float)
    assert abs(E - float(z[f"{name}/E"])) <= 1e-10
    # same vector up to a phase
    ov = torch.vdot(v, vec.cpu())
    assert abs(abs(ov) - 1.0) <= 1e-8


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.complex64, torch.complex128])
def test_against_eigh_device_dtype_matrix(dtype):                    # test_lanczos.py:127-157
    n = 24
    H = rand_herm(n, dtype, 3)
    exact = float(torch.linalg.eigvalsh(H)[0])
    g = torch.Generator().manual_seed(4)
    v0 = torch.randn(n, dtype=dtype, generator=g)
    vec, E = qsb.lanczos_ground_state(v0.to(DEV), matvec, (H.to(DEV),), num_restarts=8, krylov_dim=12, pro=True)
    assert abs(E - exact) <= atol_for(dtype) * 10
    assert abs(float(torch.linalg.norm(vec)) - 1.0) <= (1e-12 if dtype in (torch.float64, torch.complex128) else 1e-5)
    Hv = H.to(DEV) @ vec
    assert float(torch.linalg.norm(Hv - E * vec)) <= (1e-6 if dtype in (torch.float64, torch.complex128) else 5e-3)


def test_simple_hamiltonian_and_zero_start_vector():                # test_lanczos.py:81-125
    H = torch.tensor([[2.0, -1.0, 0.0, 0.0], [-1.0, 2.0, -1.0, 0.0], [0.0, -1.0, 2.0, -1.0],
                      [0.0, 0.0, -1.0, 2.0]], dtype=torch.float64)
    exact = float(torch.linalg.eigvalsh(H)[0])
    v0 = torch.ones(4, dtype=torch.float64, device=DEV)
    _, E = qsb.lanczos_ground_state(v0, matvec, (H.to(DEV),), num_restarts=3, krylov_dim=4)
    assert abs(E - exact) <= 1e-10
    torch.manual_seed(0)
    vec, E0 = qsb.lanczos_ground_state(torch.zeros(4, dtype=torch.float64, device=DEV), matvec, (H.to(DEV),),
                                       num_restarts=4, krylov_dim=4)
    assert abs(E0 - exact) <= 1e-10 and abs(float(torch.linalg.norm(vec)) - 1.0) <= 1e-12


def test_krylov_larger_than_space_breakdown():
    """krylov_dim > n: beta hits ~0; the tridiagonal matrix is truncated where the reference would stop."""
    H = rand_herm(4, torch.float64, 9)
    exact = float(torch.linalg.eigvalsh(H)[0])
    g = torch.Generator().manual_seed(1)
    v0 = torch.randn(4, dtype=torch.float64, generator=g)
    _, E = qsb.lanczos_ground_state(v0.to(DEV), matvec, (H.to(DEV),), num_restarts=2, krylov_dim=8, pro=True)
    assert abs(E - exact) <= 1e-10


def test_monotone_in_restarts_and_pro_not_worse():                  # test_lanczos.py:197-243
    H = rand_herm(32, torch.complex128, 11).to(DEV)
    g = torch.Generator().manual_seed(12)
    v0 = torch.randn(32, dtype=torch.complex128, generator=g).to(DEV)
    Es = [qsb.lanczos_ground_state(v0, matvec, (H,), num_restarts=r, krylov_dim=6)[1] for r in (1, 2, 4, 8)]
    for a, b in zip(Es, Es[1:]):
        assert b <= a + 1e-12
    E_pro = qsb.lanczos_ground_state(v0, matvec, (H,), num_restarts=4, krylov_dim=10, pro=True, pro_max_reorth=2)[1]
    E_no = qsb.lanczos_ground_state(v0, matvec, (H,), num_restarts=4, krylov_dim=10, pro=False)[1]
    exact = float(torch.linalg.eigvalsh(H.cpu())[0])
    assert abs(E_pro - exact) <= abs(E_no - exact) + 1e-10


def test_large_beta_tol_stops_early_and_residual_tol():             # test_lanczos.py:160-175, 271-290
    H = rand_herm(16, torch.float64, 13).to(DEV)
    g = torch.Generator().manual_seed(14)
    v0 = torch.randn(16, dtype=torch.float64, generator=g).to(DEV)
    v, E = qsb.lanczos_ground_state(v0, matvec, (H,), num_restarts=1, krylov_dim=8, beta_tol=1e6)
    x = v0 / torch.linalg.norm(v0)
    rq = float(torch.vdot(x, H @ x).real)
    assert abs(E - rq) <= 1e-12                     # one iteration: the Ritz value is the Rayleigh quotient
    _, E2 = qsb.lanczos_ground_state(v0, matvec, (H,), num_restarts=50, krylov_dim=10, pro=True,
                                     ritz_residual_tol=1e-9)
    assert abs(E2 - float(torch.linalg.eigvalsh(H.cpu())[0])) <= 1e-9


def test_operator_with_out_kwarg_is_used_in_place():                # test_lanczos.py:292-314
    H = rand_herm(20, torch.float64, 15).to(DEV)
    calls = {"out": 0}

    def op(x, A, out=None):
        calls["out"] += out is not None
        return torch.mv(A, x, out=out)

    g = torch.Generator().manual_seed(16)
    v0 = torch.randn(20, dtype=torch.float64, generator=g).to(DEV)
    _, E = qsb.lanczos_ground_state(v0, op, (H,), num_restarts=6, krylov_dim=8, pro=True)
    assert calls["out"] == 6 * 8
    assert abs(E - float(torch.linalg.eigvalsh(H.cpu())[0])) <= 1e-10


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128, torch.float32, torch.complex64])
@pytest.mark.parametrize("n", [1, 3, 1000, 4097, 1 << 20])
def test_vector_kernels_against_torch(dtype, n):
    """dot / axpy+norm / scale / combine against plain torch (fp64 tolerance 1e-13 relative)."""
    ops = _lib.ops
    g = torch.Generator().manual_seed(n)
    K = torch.randn(5, n, dtype=dtype, generator=g).to(DEV)
    w = torch.randn(n, dtype=dtype, generator=g).to(DEV)
    scal = torch.zeros(32, 2, dtype=torch.float64, device=DEV)
    scratch = torch.zeros(_lib.lz_scratch_bytes(), dtype=torch.uint8, device=DEV)
    tol = 1e-13 if dtype in (torch.float64, torch.complex128) else 2e-5
    ops.lz_dot(K, 5, n, w, n, scal, 0, 0, scratch)
    ref = (K.conj().to(torch.complex128) @ w.to(torch.complex128))
    got = torch.complex(scal[:5, 0], scal[:5, 1])
    assert float((got - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max())) * 10
    ops.lz_dot(w, 1, 0, w, n, scal, 8, 3, scratch)
    assert abs(float(scal[8, 0]) - float(torch.linalg.norm(w.to(torch.complex128)))) <= tol * n ** 0.5
    coef = torch.tensor([[0.3, -0.2], [1.1, 0.4], [-0.7, 0.0]], dtype=torch.float64)
    scal[10:13] = coef.to(DEV)
    w2 = w.clone()
    ops.lz_axpy_norm(K, 3, n, w2, n, scal, 10, 20, 0, scratch)
    cc = torch.complex(coef[:, 0], coef[:, 1]) if dtype.is_complex else coef[:, 0]
    refw = w.cpu().to(torch.complex128) - (cc.to(torch.complex128)[:, None] * K[:3].cpu().to(torch.complex128)).sum(0)
    assert rel_err(w2.to(torch.complex128), refw) <= tol * 10
    assert abs(float(scal[20, 0]) - float(torch.linalg.norm(refw))) <= tol * 10 * max(1.0, float(torch.linalg.norm(refw)))
    dst = torch.empty_like(w)
    ops.lz_scale(w2, dst, n, scal, 20)
    assert rel_err(dst.to(torch.complex128), refw / torch.linalg.norm(refw)) <= tol * 10
    y = [0.5, -0.25, 2.0, 0.125]
    ops.lz_combine(K, 4, n, y, dst, n, scal, 21, 0, scratch)
    refc = (torch.tensor(y, dtype=torch.complex128)[:, None] * K[:4].cpu().to(torch.complex128)).sum(0)
    assert rel_err(dst.to(torch.complex128), refc) <= tol * 10
    assert abs(float(scal[21, 0]) - float(torch.linalg.norm(refc))) <= tol * 10 * max(1.0, float(torch.linalg.norm(refc)))

"""The HOST logic of the Lanczos solver (restarts, deferred alpha/beta read-back, breakdown truncation, PRO, Ritz
update, early stop) on the CPU: eigensolver._lanczos_impl with the vector backend replaced by the test-only torch
backend (tests/vecops_torch.py) against the golden energies of the unmodified reference and against eigh.  The
product entry point lanczos_ground_state stays CUDA-only (see test_abi.test_no_cpu_fallback)."""
import pytest
import torch

from conftest import load_golden
from quantum_simulation_b200.eigensolver import _lanczos_impl
from vecops_torch import TorchVectorOps


def matvec(x, A):
    return A @ x


def solve(v0, H, **kw):
    args = dict(num_restarts=2, krylov_dim=4, pro=False, pro_max_reorth=1, beta_tol=1e-12, ritz_residual_tol=None)
    args.update(kw)
    return _lanczos_impl(v0, matvec, (H,), args["num_restarts"], args["krylov_dim"], args["pro"],
                         args["pro_max_reorth"], args["beta_tol"], args["ritz_residual_tol"], vec=TorchVectorOps())


@pytest.mark.parametrize("name", [str(s) for s in load_golden("lanczos")["names"]])
def test_host_logic_reproduces_reference_energies(name):
    z = load_golden("lanczos")
    H, v0, v = (torch.from_numpy(z[f"{name}/{k}"]) for k in ("H", "v0", "v"))
    k, r, pro, mr = (int(x) for x in z[f"{name}/params"])
    vec, E = solve(v0, H, num_restarts=r, krylov_dim=k, pro=bool(pro), pro_max_reorth=mr)
    assert abs(E - float(z[f"{name}/E"])) <= 1e-11
    assert abs(abs(torch.vdot(v, vec)) - 1.0) <= 1e-8
    assert abs(float(torch.linalg.norm(vec)) - 1.0) <= 1e-13


def rand_herm(n, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, dtype=dtype, generator=g)
    return 0.5 * (A + A.conj().T)


def test_breakdown_is_truncated_where_the_reference_stops():
    """Krylov dimension larger than the space: beta drops to ~0 at j = n-1; the deferred read-back must cut the
    tridiagonal matrix there (reference: break at lanczos_solver.py:263-265), not use the garbage that follows."""
    H = rand_herm(4, torch.float64, 2)
    exact = float(torch.linalg.eigvalsh(H)[0])
    g = torch.Generator().manual_seed(5)
    v0 = torch.randn(4, dtype=torch.float64, generator=g)
    _, E = solve(v0, H, num_restarts=2, krylov_dim=9, pro=True)
    assert abs(E - exact) <= 1e-10
    # an exact eigenvector as start vector: beta_0 = 0 -> one-dimensional Krylov space, E = its eigenvalue
    w, V = torch.linalg.eigh(H)
    _, E1 = solve(V[:, 2].clone(), H, num_restarts=1, krylov_dim=4)
    assert abs(E1 - float(w[2])) <= 1e-12


def test_zero_start_vector_restarts_and_tolerances():
    H = rand_herm(12, torch.complex128, 7)
    exact = float(torch.linalg.eigvalsh(H)[0])
    torch.manual_seed(0)
    vec, E = solve(torch.zeros(12, dtype=torch.complex128), H, num_restarts=30, krylov_dim=6, pro=True)
    assert abs(E - exact) <= 1e-10 and abs(float(torch.linalg.norm(vec)) - 1.0) <= 1e-13
    g = torch.Generator().manual_seed(8)
    v0 = torch.randn(12, dtype=torch.complex128, generator=g)
    Es = [solve(v0, H, num_restarts=r, krylov_dim=4)[1] for r in (1, 2, 4, 8)]
    assert all(b <= a + 1e-12 for a, b in zip(Es, Es[1:]))                 # monotone in the number of restarts
    x = v0 / torch.linalg.norm(v0)
    _, E_big_tol = solve(v0, H, num_restarts=1, krylov_dim=6, beta_tol=1e9)  # stops after one iteration
    assert abs(E_big_tol - float(torch.vdot(x, H @ x).real)) <= 1e-12
    _, E_res = solve(v0, H, num_restarts=100, krylov_dim=8, pro=True, ritz_residual_tol=1e-9)
    assert abs(E_res - exact) <= 1e-8
    with pytest.raises(ValueError):
        solve(v0, H, krylov_dim=65)


def test_converged_regime_stops_where_the_reference_stops():
    """after a restart that ended in a breakdown the solver watches beta every iteration (like the
    reference's per-iteration `beta.item()`, lanczos_solver.py:263) and skips the matvecs the reference skips.  An exact
    eigenvector as start: the first call still enqueues a whole restart (nothing known yet) and learns; the second call
    applies the operator once per restart -- same energies, same vectors."""
    from quantum_simulation_b200 import eigensolver
    H = rand_herm(16, torch.float64, 11)
    w, V = torch.linalg.eigh(H)
    calls = {"n": 0}

    def counted(x, A):
        calls["n"] += 1
        return A @ x
    ops = TorchVectorOps()

    def run():
        calls["n"] = 0
        v, E = _lanczos_impl(V[:, 0].clone(), counted, (H,), 2, 4, False, 1, 1e-12, None, vec=ops)
        return v.clone(), E, calls["n"]
    eigensolver.clear_workspaces()
    v1, E1, n1 = run()
    v2, E2, n2 = run()
    assert abs(E1 - float(w[0])) <= 1e-12 and E2 == E1
    assert torch.allclose(v1, v2, atol=1e-14)
    assert n1 == 4 + 1 and n2 == 2          # first restart blind (4), then watched: one matvec per restart
    # a generic start vector afterwards: beta stays large, the watch switches itself off after one full restart
    g = torch.Generator().manual_seed(3)
    calls["n"] = 0
    _lanczos_impl(torch.randn(16, dtype=torch.float64, generator=g), counted, (H,), 2, 4, False, 1, 1e-12, None, vec=ops)
    assert calls["n"] == 8

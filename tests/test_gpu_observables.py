"""GPU: the environment-based observables (quantum_simulation_b200/observables.py -- transfer-operator chains on the
qsb_env_left / qsb_env_right kernels) against the outputs of the reference's dense functions
(dmrg_observables.py:352-449, dmrg_advanced_observables.py:184-565) stored in tests/golden/observables.npz, and against
the dense oracle on states the fixture does not hold.  Tolerance 1e-11 absolute on O(1) expectation values."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import observables_oracle as oo
import quantum_simulation_b200 as qsb

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
CASES = ["heis8_f64", "rand8_c128", "rand6_spin1"]
TOL = 1e-11


def _mps(z, tag, prefix=None):
    """An MPS-like object (the observables only read Nsites / chid / A / B / sWeight)."""
    pre = tag if prefix is None else prefix
    N = 8 if prefix == "fid_b" else int(z[f"{tag}/cfg"][0])
    B = [torch.from_numpy(z[f"{pre}/B{p}"]).to(DEV) for p in range(N)]
    m = types.SimpleNamespace(Nsites=N, chid=B[0].shape[1], B=B, A=B)
    if prefix is None:
        m.sWeight = [torch.from_numpy(z[f"{tag}/s{p}"]).to(DEV) for p in range(N + 1)]
    return m


def _close(got, ref, tol=TOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert np.allclose(got, ref, rtol=tol, atol=tol), float(np.max(np.abs(got - ref)))


@pytest.mark.parametrize("tag", CASES)
def test_local_two_point_string_match_reference(tag):
    z = load_golden("observables")
    m = _mps(z, tag)
    N, d = m.Nsites, m.chid
    names = ["sz", "sx", "sy", "num"] if d == 2 else ["s1z", "s1str"]
    for name in names:
        O = torch.from_numpy(z[f"{tag}/op_{name}"])
        got = torch.stack([qsb.local_expectation(m, O, i) for i in range(N)])
        _close(got, z[f"{tag}/local_{name}"])
        # dtype promotion as in the reference: a complex operator on a real state gives a complex value
        assert got.dtype.is_complex == (O.dtype.is_complex or m.B[0].dtype.is_complex)
    a, b = names[0], names[1]
    Oa, Ob = torch.from_numpy(z[f"{tag}/op_{a}"]), torch.from_numpy(z[f"{tag}/op_{b}"])
    for (i, j), ref in zip(z[f"{tag}/two_point_pairs"], z[f"{tag}/two_point_{a}_{b}"]):
        _close(qsb.two_point_correlation(m, Oa, int(i), Ob, int(j)), ref)
    L_, S_, R_ = ((z[f"{tag}/op_sz"], 2 * z[f"{tag}/op_sz"], z[f"{tag}/op_sz"]) if d == 2 else
                  (z[f"{tag}/op_s1z"], z[f"{tag}/op_s1str"], z[f"{tag}/op_s1z"]))
    for (i, j), ref in zip(z[f"{tag}/string_sites"], z[f"{tag}/string_values"]):
        _close(qsb.string_order_parameter(m, torch.from_numpy(L_), torch.from_numpy(S_), torch.from_numpy(R_), int(i), int(j)),
               ref)
    with pytest.raises(ValueError):
        qsb.local_expectation(m, Oa, N)
    with pytest.raises(ValueError):
        qsb.local_expectation(m, torch.eye(d + 1), 0)
    with pytest.raises(ValueError):
        qsb.string_order_parameter(m, Oa, Oa, Oa, 3, 3)


@pytest.mark.parametrize("tag", CASES)
def test_structure_factor_polarization_entropy_match_reference(tag):
    z = load_golden("observables")
    m = _mps(z, tag)
    d = m.chid
    for name in (["sz", "sy"] if d == 2 else ["s1z"]):
        O = torch.from_numpy(z[f"{tag}/op_{name}"])
        for conn in (0, 1):
            sf = qsb.structure_factor(m, O, connected=bool(conn))
            _close(sf["correlation_matrix"], z[f"{tag}/sf_{name}_{conn}_corr"])
            _close(sf["values"], z[f"{tag}/sf_{name}_{conn}_values"], 1e-10)
        _close(qsb.structure_factor(m, O, q_values=[0.3, 1.7])["values"], z[f"{tag}/sfq_{name}_values"], 1e-10)
    if d == 2:
        for k, origin in enumerate((0, 1)):
            p = qsb.many_body_polarization(m, torch.from_numpy(z[f"{tag}/op_num"]), origin=origin)
            ref = z[f"{tag}/polarization"][k]
            _close(p["expectation"], ref[0])
            _close(p["phase"], ref[1].real, 1e-9)
            _close(p["polarization"], ref[2].real, 1e-9)
    _close(np.array(qsb.entanglement_entropy(m)), z[f"{tag}/entropy"], 1e-12)
    _close(qsb.entanglement_spectrum(m, m.Nsites // 2), z[f"{tag}/spectrum_mid"], 1e-12)
    key = f"{tag}/corrlen"
    first = "sz" if d == 2 else "s1z"
    if key in z.files:
        cl = qsb.correlation_length(m, torch.from_numpy(z[f"{tag}/op_{first}"]), connected=True, fit_range=(1, 3))
        _close(cl["correlations"], z[f"{tag}/corrlen_corr"])
        ref = z[key]
        assert abs(cl["xi"] - ref[0]) <= 1e-7 * abs(ref[0]) and abs(cl["slope"] - ref[1]) <= 1e-8
    else:
        with pytest.raises(ValueError):
            qsb.correlation_length(m, torch.from_numpy(z[f"{tag}/op_{first}"]), connected=True, fit_range=(1, 3))


def test_fidelity_between_mps_of_different_bond_dimensions():
    z = load_golden("observables")
    a, c = _mps(z, "heis8_f64"), _mps(z, "rand8_c128")
    b = _mps(z, "heis8_f64", prefix="fid_b")
    _close(qsb.ground_state_fidelity(a, b), z["fidelity"][0])
    _close(qsb.ground_state_fidelity(a, c), z["fidelity"][1])
    _close(qsb.ground_state_fidelity(a, a), 1.0)
    _close(qsb.ground_state_fidelity(c, a), z["fidelity"][1])


def test_observables_on_a_state_the_dense_path_cannot_hold():
    """L = 40, chi = 64 (d^N = 1e12 amplitudes): sum rules instead of a dense check.  For a normalised state
    sum_ab <S^a_i S^a_j> over a = x, y, z at i = j is S(S+1) = 3/4, <n> of the number operator stays in [0, 1], the
    fidelity of a state with itself is 1, and the structure factor's correlation matrix is symmetric with the local
    values on ... its connected diagonal <O^2> - <O>^2."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = qsb.MPS(40, 2, 64, DEV, torch.complex128, seed=4)
    sx = torch.tensor([[0.0, 0.5], [0.5, 0.0]], dtype=torch.float64)
    sy = torch.tensor([[0.0, -0.5j], [0.5j, 0.0]], dtype=torch.complex128)
    sz = torch.tensor([[0.5, 0.0], [0.0, -0.5]], dtype=torch.float64)
    tot = sum(qsb.two_point_correlation(m, o, 17, o, 17) for o in (sx, sy, sz))
    assert abs(complex(tot) - 0.75) <= 1e-12
    sf = qsb.structure_factor(m, sz, connected=True)
    C = sf["correlation_matrix"]
    loc = torch.stack([qsb.local_expectation(m, sz, i) for i in range(40)])
    assert float((C - C.T).abs().max()) <= 1e-13
    assert float((torch.diagonal(C) - (0.25 - loc ** 2)).abs().max()) <= 1e-12
    assert abs(float(qsb.ground_state_fidelity(m, m)) - 1.0) <= 1e-12
    # oracle cross-check on a sub-chain small enough to be dense: the first 10 sites are not a state of their own, so
    # compare a full small MPS instead
    with contextlib.redirect_stdout(io.StringIO()):
        s = qsb.MPS(10, 2, 16, DEV, torch.float64, seed=8)
    T = [t.detach().cpu().numpy() for t in s.B]
    for (i, j) in [(0, 9), (3, 4), (7, 2)]:
        ref = oo.two_point_correlation(T, sy.numpy(), i, sx.numpy(), j)
        assert abs(complex(qsb.two_point_correlation(s, sy, i, sx, j)) - ref) <= 1e-12

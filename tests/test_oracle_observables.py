"""CPU: the dense observables oracle (oracle/observables_oracle.py) against the fixture written by the UNMODIFIED
reference (tests/golden/observables.npz, generator: tests/golden/make_golden_observables.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import observables_oracle as oo

CASES = ["heis8_f64", "rand8_c128", "rand6_spin1"]


def _tensors(z, tag):
    N = int(z[f"{tag}/cfg"][0])
    return [z[f"{tag}/B{p}"] for p in range(N)]


@pytest.mark.parametrize("tag", CASES)
def test_oracle_local_two_point_string(tag):
    z = load_golden("observables")
    T = _tensors(z, tag)
    N, d = (int(x) for x in z[f"{tag}/cfg"])
    names = ["sz", "sx", "sy", "num"] if d == 2 else ["s1z", "s1str"]
    for name in names:
        O = z[f"{tag}/op_{name}"]
        got = np.array([oo.local_expectation(T, O, i) for i in range(N)])
        assert np.allclose(got, z[f"{tag}/local_{name}"], rtol=1e-11, atol=1e-12)
    a, b = names[0], names[1]
    for (i, j), ref in zip(z[f"{tag}/two_point_pairs"], z[f"{tag}/two_point_{a}_{b}"]):
        v = oo.two_point_correlation(T, z[f"{tag}/op_{a}"], int(i), z[f"{tag}/op_{b}"], int(j))
        assert abs(v - ref) <= 1e-12
    L_, S_, R_ = ((z[f"{tag}/op_sz"], 2 * z[f"{tag}/op_sz"], z[f"{tag}/op_sz"]) if d == 2 else
                  (z[f"{tag}/op_s1z"], z[f"{tag}/op_s1str"], z[f"{tag}/op_s1z"]))
    for (i, j), ref in zip(z[f"{tag}/string_sites"], z[f"{tag}/string_values"]):
        assert abs(oo.string_order_parameter(T, L_, S_, R_, int(i), int(j)) - ref) <= 1e-12


@pytest.mark.parametrize("tag", CASES)
def test_oracle_structure_factor_and_polarization(tag):
    z = load_golden("observables")
    T = _tensors(z, tag)
    d = int(z[f"{tag}/cfg"][1])
    name = "sz" if d == 2 else "s1z"
    for conn in (0, 1):
        o = oo.structure_factor(T, z[f"{tag}/op_{name}"], connected=bool(conn))
        assert np.allclose(o["values"], z[f"{tag}/sf_{name}_{conn}_values"], rtol=1e-10, atol=1e-12)
        assert np.allclose(o["correlation_matrix"], z[f"{tag}/sf_{name}_{conn}_corr"], rtol=1e-10, atol=1e-12)
    o = oo.structure_factor(T, z[f"{tag}/op_{name}"], q_values=[0.3, 1.7])
    assert np.allclose(o["values"], z[f"{tag}/sfq_{name}_values"], rtol=1e-10, atol=1e-12)
    if d == 2:
        for k, origin in enumerate((0, 1)):
            p = oo.many_body_polarization(T, z[f"{tag}/op_num"], origin)
            assert abs(p["expectation"] - z[f"{tag}/polarization"][k, 0]) <= 1e-12


def test_oracle_fidelity():
    z = load_golden("observables")
    a, c = _tensors(z, "heis8_f64"), _tensors(z, "rand8_c128")
    b = [z[f"fid_b/B{p}"] for p in range(8)]
    assert abs(oo.fidelity(a, b) - z["fidelity"][0]) <= 1e-12
    assert abs(oo.fidelity(a, c) - z["fidelity"][1]) <= 1e-12
    assert abs(z["fidelity"][2] - 1.0) <= 1e-12

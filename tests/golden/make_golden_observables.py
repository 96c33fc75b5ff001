#!/usr/bin/env python3
"""Fixture for the environment-based observables: outputs of the UNMODIFIED reference's
dense functions (dmrg_observables.py:352-449, dmrg_advanced_observables.py:184-565) on small converged / random MPS.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_observables.py

Also asserts that oracle/observables_oracle.py reproduces every stored value (the oracle's pin)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg      # noqa: E402  (import stubs + reference on sys.path)
from quantum_simulation import dmrg_observables as ro          # noqa: E402
from quantum_simulation import dmrg_advanced_observables as ra  # noqa: E402
from oracle import observables_oracle as oo                    # noqa: E402

F64, C128 = torch.float64, torch.complex128
SZ = torch.tensor([[0.5, 0.0], [0.0, -0.5]], dtype=F64)
SX = torch.tensor([[0.0, 0.5], [0.5, 0.0]], dtype=F64)
SY = torch.tensor([[0.0, -0.5j], [0.5j, 0.0]], dtype=C128)
NUM = torch.tensor([[1.0, 0.0], [0.0, 0.0]], dtype=F64)
S1Z = torch.diag(torch.tensor([1.0, 0.0, -1.0], dtype=F64))
S1STR = torch.diag(torch.exp(1j * np.pi * torch.tensor([1.0, 0.0, -1.0]))).real.to(F64)


def converged(mpo, L, d, chi, dt, seed, sweeps=3):
    mps = mg.quiet(mg.MPS, L, d, chi, "cpu", dt, seed)
    sw = mg.Sweeps(sweeps)
    sw.set_schedule(maxdim=[chi] * sweeps, cutoff=[1e-12] * sweeps, krylov_dim=[4] * sweeps, noise=[0.0] * sweeps)
    dm = mg.quiet(mg.DMRG, mps, mpo, device="cpu", dtype=dt)
    mg.quiet(dm, sw, dispon=0)
    return dm.mps


def close(a, b, tol=1e-11):
    a, b = np.asarray(a), np.asarray(b)
    assert np.allclose(a, b, rtol=tol, atol=tol), (a, b)


store = {}
cases = {
    "heis8_f64": (converged(mg.quiet(mg.heisenberg_real_mpo, 8), 8, 2, 12, F64, 3), 2),
    "rand8_c128": (mg.quiet(mg.MPS, 8, 2, 6, "cpu", C128, 5), 2),
    "rand6_spin1": (mg.quiet(mg.MPS, 6, 3, 7, "cpu", F64, 9), 3),
}
for tag, (mps, d) in cases.items():
    N = mps.Nsites
    T = [t.detach().numpy() for t in mps.B]
    for p, t in enumerate(T):
        store[f"{tag}/B{p}"] = t
    for p in range(N + 1):
        store[f"{tag}/s{p}"] = mps.sWeight[p].detach().numpy()
    store[f"{tag}/cfg"] = np.array([N, d])
    ops = {"sz": SZ, "sx": SX, "sy": SY, "num": NUM} if d == 2 else {"s1z": S1Z, "s1str": S1STR}
    for name, O in ops.items():
        store[f"{tag}/op_{name}"] = O.numpy()
        loc = np.array([ro.local_expectation(mps, O, i).numpy() for i in range(N)])
        store[f"{tag}/local_{name}"] = loc
        close(loc, [oo.local_expectation(T, O.numpy(), i) for i in range(N)])
        for conn in (False, True):
            sf = ra.structure_factor(mps, O, connected=conn)
            store[f"{tag}/sf_{name}_{int(conn)}_values"] = sf["values"].numpy()
            store[f"{tag}/sf_{name}_{int(conn)}_corr"] = sf["correlation_matrix"].numpy()
            o = oo.structure_factor(T, O.numpy(), connected=conn)
            close(sf["values"].numpy(), o["values"])
            close(sf["correlation_matrix"].numpy(), o["correlation_matrix"])
        sfq = ra.structure_factor(mps, O, q_values=[0.3, 1.7])
        store[f"{tag}/sfq_{name}_values"] = sfq["values"].numpy()
    names = list(ops)
    pairs = [(0, N - 1), (2, 5 if N > 5 else N - 1), (4, 1), (3, 3)]
    tp = []
    for (i, j) in pairs:
        v = ro.two_point_correlation(mps, ops[names[0]], i, ops[names[1]], j)
        tp.append(complex(v))
        close(complex(v), oo.two_point_correlation(T, ops[names[0]].numpy(), i, ops[names[1]].numpy(), j))
    store[f"{tag}/two_point_pairs"] = np.array(pairs)
    store[f"{tag}/two_point_{names[0]}_{names[1]}"] = np.array(tp)
    if d == 2:
        so = [(0, 3), (1, 2), (2, N - 1)]
        vals = [complex(ra.string_order_parameter(mps, SZ, 2.0 * SZ, SZ, i, j)) for i, j in so]
        pol = [ra.many_body_polarization(mps, NUM, origin=o) for o in (0, 1)]
        for o_, pp in zip((0, 1), pol):
            close(complex(pp["expectation"]), oo.many_body_polarization(T, NUM.numpy(), o_)["expectation"])
        store[f"{tag}/polarization"] = np.array([[complex(p["expectation"]), float(p["phase"]), float(p["polarization"])]
                                                 for p in pol])
    else:
        so = [(0, 2), (1, 5), (2, 3)]
        vals = [complex(ra.string_order_parameter(mps, S1Z, S1STR, S1Z, i, j)) for i, j in so]
    for (i, j), v in zip(so, vals):
        L_, S_, R_ = ((SZ, 2.0 * SZ, SZ) if d == 2 else (S1Z, S1STR, S1Z))
        close(v, oo.string_order_parameter(T, L_.numpy(), S_.numpy(), R_.numpy(), i, j))
    store[f"{tag}/string_sites"] = np.array(so)
    store[f"{tag}/string_values"] = np.array(vals)
    store[f"{tag}/entropy"] = np.array(ro.entanglement_entropy(mps))
    spec = ra.entanglement_spectrum(mps)
    store[f"{tag}/spectrum_mid"] = spec[N // 2].numpy()
    try:
        first = names[0]
        cl = ra.correlation_length(mps, ops[first], connected=True, fit_range=(1, 3))
        store[f"{tag}/corrlen"] = np.array([cl["xi"], cl["slope"], cl["intercept"]])
        store[f"{tag}/corrlen_corr"] = cl["correlations"].numpy()
    except ValueError as ex:
        store[f"{tag}/corrlen_error"] = np.array(str(ex))
# fidelity between MPS of different bond dimensions
a = cases["heis8_f64"][0]
b = mg.quiet(mg.MPS, 8, 2, 5, "cpu", F64, 21)
c = cases["rand8_c128"][0]
for p, t in enumerate(b.B):
    store[f"fid_b/B{p}"] = t.detach().numpy()
store["fidelity"] = np.array([float(ra.ground_state_fidelity(a, b)), float(ra.ground_state_fidelity(a, c)),
                              float(ra.ground_state_fidelity(a, a))])
close(store["fidelity"][0], oo.fidelity([t.detach().numpy() for t in a.B], [t.detach().numpy() for t in b.B]))
close(store["fidelity"][1], oo.fidelity([t.detach().numpy() for t in a.B], [t.detach().numpy() for t in c.B]))
np.savez_compressed(os.path.join(HERE, "observables.npz"), **store)
print("wrote observables.npz;", len(store), "arrays; oracle reproduces every value")

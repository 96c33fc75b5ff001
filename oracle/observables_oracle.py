"""TEST INFRASTRUCTURE ONLY -- dense CPU restatement of the reference's product-operator observables.

Only tests/ may import this module; the product path (quantum-simulation_b200/observables.py) contracts transfer
operators on the CUDA environment kernels and never touches it.  Every function follows the reference's dense
definition: the MPS is contracted to a d^N vector (dmrg_observables.py:236-263), local operators are applied site by
site (:307-330) and expectation values are <psi|O|psi>/<psi|psi> (:333-349).  Pinned by tests/golden/observables.npz,
which tests/golden/make_golden_observables.py generates by calling the UNMODIFIED reference functions.
"""
import numpy as np


def dense_state(tensors):
    """Product of the site tensors (l, d, r), flattened (dmrg_observables.py:258-263)."""
    state = np.asarray(tensors[0])
    for t in tensors[1:]:
        state = np.tensordot(state, np.asarray(t), axes=([-1], [0]))
    return state.reshape(-1)


def apply_local(state, op, site, n, d):
    """Operator on one site of the dense vector (dmrg_observables.py:307-330)."""
    psi = state.reshape((d,) * n)
    psi = np.moveaxis(np.tensordot(np.asarray(op), psi, axes=([1], [site])), 0, site)
    return psi.reshape(-1)


def expect(state, applied):
    nrm = np.vdot(state, state)
    if abs(nrm) == 0:
        raise ValueError("zero-norm state")
    return np.vdot(state, applied) / nrm


def product_expectation(tensors, ops):
    """< prod_k O_k > for a dict site -> operator."""
    n, d = len(tensors), tensors[0].shape[1]
    dt = np.result_type(tensors[0].dtype, *[np.asarray(o).dtype for o in ops.values()])
    state = dense_state(tensors).astype(dt)
    applied = state
    for site in sorted(ops, reverse=True):
        applied = apply_local(applied, np.asarray(ops[site]).astype(dt), site, n, d)
    return expect(state, applied)


def local_expectation(tensors, op, site):                              # dmrg_observables.py:352-389
    return product_expectation(tensors, {site: op})


def two_point_correlation(tensors, op_i, i, op_j, j):                  # dmrg_observables.py:392-449
    if i == j:
        return local_expectation(tensors, np.asarray(op_i) @ np.asarray(op_j), i)
    return product_expectation(tensors, {i: op_i, j: op_j})


def string_order_parameter(tensors, left, string, right, i, j):        # dmrg_advanced_observables.py:417-468
    ops = {i: left, j: right}
    for k in range(i + 1, j):
        ops[k] = string
    return product_expectation(tensors, ops)


def _expm(a):
    w, v = np.linalg.eig(a)
    return (v * np.exp(w)) @ np.linalg.inv(v)


def many_body_polarization(tensors, number, origin=0):                 # dmrg_advanced_observables.py:506-565
    n = len(tensors)
    ops = {s: _expm((2j * np.pi * float(s + origin) / n) * np.asarray(number, dtype=np.complex128)) for s in range(n)}
    e = product_expectation(tensors, ops)
    return {"expectation": e, "phase": np.angle(e), "polarization": np.angle(e) / (2 * np.pi)}


def correlation_matrix(tensors, op, connected=False):                  # dmrg_advanced_observables.py:221-247
    n = len(tensors)
    loc = np.array([local_expectation(tensors, op, i) for i in range(n)])
    c = np.array([[two_point_correlation(tensors, op, i, op, j) for j in range(n)] for i in range(n)])
    return c - loc[:, None] * loc[None, :] if connected else c


def structure_factor(tensors, op, q_values=None, connected=False):     # dmrg_advanced_observables.py:184-261
    n = len(tensors)
    c = correlation_matrix(tensors, op, connected)
    q = 2 * np.pi * np.arange(n) / n if q_values is None else np.asarray(q_values, dtype=np.float64).reshape(-1)
    pos = np.arange(n, dtype=np.float64)
    sep = pos[:, None] - pos[None, :]
    vals = []
    for qq in q:
        ph = np.exp(1j * qq * sep)
        ph = ph if np.iscomplexobj(c) else ph.real                     # the reference casts the phase to corr's dtype
        vals.append((ph * c).sum() / n)
    return {"q_values": q, "values": np.array(vals), "correlation_matrix": c}


def fidelity(ta, tb):                                                  # dmrg_advanced_observables.py:471-503
    a, b = dense_state(ta), dense_state(tb)
    return abs(np.vdot(a / np.linalg.norm(a), b / np.linalg.norm(b)))

#!/usr/bin/env python3
"""Full-DMRG fixtures for the MODEL FAMILIES of BASELINE configs 4 and 5 at sizes the CPU reference can afford:
  c4s  J1-J2 Heisenberg cylinder 4 x 4 (16 sites, periodic in y, J2/J1 = 0.5): wide FSM MPO bonds, float64
  c5s  Fermi-Hubbard chain via Jordan-Wigner, L = 10, d = 4 (U = 4): complex128 like config 5
run by the UNMODIFIED reference ('ncon' branch), with the oracle asserted against it (per-update energies 1e-13).
The MPO tensors are stored in the same file (mpo.npz only holds the full-size chains).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_c45.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg      # noqa: E402

cases = [
    ("c4s_j1j2_4x4_f64", lambda: mg.quiet(mg.j1j2_cylinder_mpo, 4, 4), 16, 2, 48, mg.F64, 2,
     dict(numsweeps=3, maxdim=[48, 48, 64], krylov_dim=[4, 4, 5], cutoff=[0.0, 0.0, 1e-12], reortho=[False, False, True],
          noise=[0.0] * 3)),
    ("c5s_hubbard_L10_c128", lambda: mg.quiet(mg.hubbard_jw_mpo, 10), 10, 4, 40, mg.C128, 4,
     dict(numsweeps=3, maxdim=[40, 40, 56], krylov_dim=[4, 4, 5], cutoff=[0.0, 0.0, 1e-12], reortho=[False, False, False],
          noise=[0.0] * 3)),
]
store, names = {}, []
for name, build, L, d, chi0, dt, seed, sched in cases:
    mpo = [m.detach().to(dt) for m in build()]
    E, h, s0 = mg.run_ref_dmrg(mpo, L, d, chi0, dt, seed, sched, "ncon")
    Eo, ho, s0o = mg.run_orc_dmrg(mpo, L, d, chi0, dt, seed, dict(sched))
    assert np.array_equal(s0, s0o) and ho["bond_dims"] == h["bond_dims"], name
    rel = np.max(np.abs(E - Eo) / np.abs(E))
    assert rel < 1e-12, (name, rel)
    mg.pack_mpo(store, f"{name}/mpo", mpo)
    store[f"{name}/energies"] = E
    store[f"{name}/bond_dims"] = np.array(h["bond_dims"])
    store[f"{name}/discarded"] = np.array(h["discarded_weights"])
    store[f"{name}/init_abs_sums"] = s0
    store[f"{name}/cfg"] = np.array([L, d, chi0, seed, sched["numsweeps"]])
    store[f"{name}/dtype"] = np.array("c128" if dt == mg.C128 else "f64")
    for k in ("maxdim", "krylov_dim", "cutoff", "reortho"):
        store[f"{name}/{k}"] = np.array(sched[k])
    store[f"{name}/mpo_bonds"] = np.array([m.shape[1] for m in mpo])
    names.append(name)
    print(f"{name}: E_final = {E[-1]:.12f}, {len(E)} updates, max MPO bond {max(m.shape[1] for m in mpo)}, "
          f"max bond dim {max(h['bond_dims'])}, oracle rel diff {rel:.1e}")
store["names"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "dmrg_c45.npz"), **store)
print("wrote dmrg_c45.npz")

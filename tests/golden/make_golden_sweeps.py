#!/usr/bin/env python3
"""Multi-sweep (converged) DMRG fixtures at BASELINE-size chains, run by the UNMODIFIED reference on the CPU.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sweeps.py c2      # TFIM L=100 chi=512 c128, 10 sweeps
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sweeps.py c3s     # Heisenberg L=100 chi=256 f64, 4 sweeps

Why these exist ("Next round" 1): from a RANDOM start the first sweep is chaotic even for the reference
against itself (its 'ncon' and 'einsum' branches differ by 1e-7 after one sweep), but the variational optimum at a
fixed bond dimension is unique, so end-of-sweep energies of LATER sweeps are reproducible to 1e-10 relative -- the bar
BASELINE.json states ("ground-state energy per sweep within 1e-10 relative").  Each fixture stores every local-update
energy, the energies at the end of each sweep, and -- as the reference's own noise floor -- the same run through the
reference's second contraction branch ('einsum') when asked for (`--both`).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg      # noqa: E402  (installs the import stubs and loads the reference)

CASES = {
    # name: (mpo builder, L, d, chi, dtype, seed, sweeps)
    "c2": (lambda: mg.quiet(mg.Ising1DMPOBuilder, 100, 1.0, 0.5, dtype=mg.C128).get_mpo(), 100, 2, 512, mg.C128, 1, 10),
    "c3s": (lambda: mg.quiet(mg.heisenberg_real_mpo, 100), 100, 2, 256, mg.F64, 1, 4),
    "c3s128": (lambda: mg.quiet(mg.heisenberg_real_mpo, 100), 100, 2, 128, mg.F64, 1, 6),
}

if __name__ == "__main__":
    name = sys.argv[1]
    both = "--both" in sys.argv
    build, L, d, chi, dt, seed, ns = CASES[name]
    for a in sys.argv[2:]:
        if a.startswith("--sweeps="):
            ns = int(a.split("=")[1])
    mpo = [t.detach() for t in build()]
    sched = dict(numsweeps=ns, maxdim=[chi] * ns, krylov_dim=[4] * ns, cutoff=[0.0] * ns, reortho=[False] * ns,
                 noise=[0.0] * ns)
    out = {}
    for be in (("ncon", "einsum") if both else ("ncon",)):
        t0 = time.time()
        E, h, s0 = mg.run_ref_dmrg(mpo, L, d, chi, dt, seed, sched, be)
        dt_s = time.time() - t0
        per = len(E) // ns
        ends = E[per - 1::per]
        print(f"reference[{be}]: {dt_s:.1f} s; end-of-sweep energies:", flush=True)
        for k, e in enumerate(ends):
            print(f"   sweep {k + 1}: {e:.15f}", flush=True)
        out[be] = (E, h, s0, dt_s)
    E, h, s0, secs = out["ncon"]
    per = len(E) // ns
    store = dict(energies=E, sweep_end=E[per - 1::per], bond_dims=np.array(h['bond_dims']),
                 discarded=np.array(h['discarded_weights']), init_abs_sums=s0,
                 cfg=np.array([L, d, chi, seed, ns]), dtype=np.array("c128" if dt == mg.C128 else "f64"),
                 ref_seconds=np.array(secs), threads=np.array(torch.get_num_threads()))
    if both:
        Ee = out["einsum"][0]
        store["einsum_energies"] = Ee
        store["einsum_sweep_end"] = Ee[per - 1::per]
        rel = np.abs(E[per - 1::per] - Ee[per - 1::per]) / np.abs(E[per - 1::per])
        print("ncon vs einsum, relative end-of-sweep deviation per sweep:", " ".join(f"{x:.1e}" for x in rel))
    np.savez_compressed(os.path.join(HERE, f"dmrg_sweeps_{name}.npz"), **store)
    print("done")

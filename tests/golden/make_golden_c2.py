#!/usr/bin/env python3
"""BASELINE config 2 fixture: TFIM L=100, g=0.5, chi=512, complex128 -- the FIRST sweep (198 local updates) run by
the UNMODIFIED reference on the CPU (ncon backend), with the oracle asserted against it.  Takes tens of minutes on
8 cores; kept separate from make_golden.py.   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_c2.py"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg      # noqa: E402  (installs the import stubs and loads the reference)

if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    mpo = mg.quiet(mg.Ising1DMPOBuilder, 100, 1.0, 0.5, dtype=mg.C128).get_mpo()
    mpo = [t.detach() for t in mpo]
    sched = dict(numsweeps=1, maxdim=[chi], krylov_dim=[4], cutoff=[0.0], reortho=[False], noise=[0.0])
    t0 = time.time()
    E, h, s0 = mg.run_ref_dmrg(mpo, 100, 2, chi, mg.C128, 1, sched, 'ncon')
    t_ref = time.time() - t0
    print(f"reference: {t_ref:.1f} s, E_end = {E[-1]:.12f}", flush=True)
    t0 = time.time()
    Eo, ho, s0o = mg.run_orc_dmrg(mpo, 100, 2, chi, mg.C128, 1, sched)
    t_orc = time.time() - t0
    rel = float(np.max(np.abs(E - Eo) / np.abs(E)))
    print(f"oracle: {t_orc:.1f} s, max rel dev vs reference {rel:.2e}", flush=True)
    assert ho['bond_dims'] == h['bond_dims']
    np.savez_compressed(os.path.join(HERE, f"dmrg_c2_chi{chi}.npz"), energies=E, oracle_energies=Eo,
                        bond_dims=np.array(h['bond_dims']), discarded=np.array(h['discarded_weights']),
                        init_abs_sums=s0, cfg=np.array([100, 2, chi, 1, 1]), ref_seconds=np.array(t_ref),
                        oracle_seconds=np.array(t_orc), threads=np.array(torch.get_num_threads()))
    print("done")

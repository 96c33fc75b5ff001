#!/usr/bin/env python3
"""Evidence for the C3 drift: the UNMODIFIED reference's MPS constructor at L=100, chi=2048.
The un-normalised carry of its right-to-left QR sweep grows by ~0.5*sqrt(chi_l*d*chi_r) per site (uniform[0,1) entries
have a rank-1 mean component), reaches ~1e280 at site 0, `torch.linalg.norm` overflows to inf and site 0 becomes
exactly zero -- so the first Lanczos start vector is zero and `_normalize_to` (lanczos_solver.py:150-153) re-seeds it
from the GLOBAL (unseeded) RNG.  Runs on CPU, needs /root/reference (container only)."""
import contextlib, io, json, sys, time
import torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import make_golden as mg          # installs the import stubs (matplotlib ...) and loads the unmodified reference
MPS = mg.MPS
t0 = time.time()
with contextlib.redirect_stdout(io.StringIO()):
    m = MPS(100, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, "cpu", torch.float64, seed=1)
a0 = m.A[0]
print(json.dumps({"reference_mps": "L=100 d=2 chi=%s f64 cpu seed=1" % (sys.argv[1] if len(sys.argv) > 1 else 2048),
                  "site0_absmax": float(a0.abs().max()), "site0_all_zero": bool((a0 == 0).all()),
                  "site1_absmax": float(m.A[1].abs().max()), "seconds": time.time() - t0}))

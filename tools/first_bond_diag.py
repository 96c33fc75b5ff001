#!/usr/bin/env python3
"""Bisect the run-to-run drift of the seeded C3 sweep down to one call.

tools/determinism_check.py localised the first differing intermediate of two identical L=100, chi=2048 sweeps to the
FIRST Lanczos result (record 100: 99 environment updates and theta0 are bitwise equal).  This tool rebuilds exactly
that state and repeats the first bond's calls, hashing the matvec output, every alpha/beta and the Ritz vector.

    python tools/first_bond_diag.py [--chi 2048] [--L 100] [--bond 0]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import quantum_simulation_b200 as qsb           # noqa: E402
from quantum_simulation_b200 import _lib, eigensolver  # noqa: E402
from determinism_check import h, mpo, quiet, DEV  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    dt = torch.float64
    Ms = mpo("heis_f64_L100", dt)
    Ms = Ms[:args.L // 2] + Ms[-(args.L - args.L // 2):]
    mps = quiet(qsb.MPS, args.L, 2, args.chi, DEV, dt, 1).to(DEV)
    env = qsb.EnvironmentManager(Ms, DEV, dt)
    for p in range(args.L - 1, 0, -1):
        env.update_R(p, mps.B[p])
    op = qsb.ApplyMPO()
    for p in (0, 1, 2):
        if p > 0:
            env.update_L(p - 1, mps.B[p - 1])         # not the sweep's A, good enough to have an L of the right shape
        Bl, Br = mps.B[p], mps.B[p + 1]
        l, s, m = Bl.shape
        _, t, r = Br.shape
        psi = (Bl.reshape(l * s, m) @ Br.reshape(m, t * r)).reshape(-1)
        L, R = env.get_L(p), env.get_R(p + 2)
        info = {"bond": p, "n": psi.numel(), "L": list(L.shape), "R": list(R.shape),
                "identL": _lib.identity_of(L), "identR": _lib.identity_of(R), "psi": h(psi), "Lh": h(L), "Rh": h(R),
                "R_absmax_per_w": [float(x) for x in R.abs().amax(dim=(1, 2))],
                "L_absmax_per_w": [float(x) for x in L.abs().amax(dim=(1, 2))]}
        print(json.dumps(info), flush=True)
        mv = [h(op(psi, L, Ms[p], Ms[p + 1], R).clone()) for _ in range(args.reps)]
        print(json.dumps({"bond": p, "matvec_hashes": mv, "repeatable": len(set(mv)) == 1}), flush=True)
        # dense H_eff (n <= 64): apply to the unit vectors
        n = psi.numel()
        if n <= 256:
            eye = torch.eye(n, dtype=dt, device=DEV)
            H = torch.stack([op(eye[i].contiguous(), L, Ms[p], Ms[p + 1], R).clone() for i in range(n)], dim=1)
            ev = torch.linalg.eigvalsh(0.5 * (H + H.T))
            print(json.dumps({"bond": p, "H_asym": float((H - H.T).abs().max()), "H_absmax": float(H.abs().max()),
                              "evals_low": [float(x) for x in ev[:6]], "evals_high": [float(x) for x in ev[-3:]]}), flush=True)
        lz = []
        for _ in range(args.reps):
            v, e = qsb.lanczos_ground_state(psi, op, (L, Ms[p], Ms[p + 1], R), num_restarts=2, krylov_dim=4)
            ws = next(iter(eigensolver._workspaces.values())) if len(eigensolver._workspaces) == 1 else None
            key = [k for k in eigensolver._workspaces if k[0] == n][0]
            ws = eigensolver._workspaces[key]
            lz.append({"e": e, "vec": h(v), "scal": [float(x) for x in ws.scal[:9, 0].cpu()], "K": h(ws.K)})
        print(json.dumps({"bond": p, "lanczos": lz, "repeatable": len({x["vec"] for x in lz}) == 1}), flush=True)


if __name__ == "__main__":
    main()

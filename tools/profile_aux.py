#!/usr/bin/env python3
"""Small driver for ncu captures of the HBM-bound kernels at BASELINE config 3 size (chi=2048, f64):
one Lanczos restart of 2 Krylov iterations with PRO (dot, axpy+norm, scale, combine) around the real matvec
(mix_prepare, mix_apply, the two GEMMs).   ncu -k regex:'lz_|mix_apply' python tools/profile_aux.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import quantum_simulation_b200 as qsb           # noqa: E402

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L, M1, M2, R, th = bench.make_inputs(chi, torch.float64, torch.device("cuda"))
L = L + L.transpose(1, 2)
R = R + R.transpose(1, 2)
_, E = qsb.lanczos_ground_state(th, qsb.ApplyMPO(), (L, M1, M2, R), num_restarts=1, krylov_dim=3, pro=True)
torch.cuda.synchronize()
print("E =", E)

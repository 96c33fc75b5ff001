#!/usr/bin/env python3
"""Fixture for the MPS mirror: tensors of the UNMODIFIED reference MPS after construction and after a sequence of
position() moves (mps.py:21-145).   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mps.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg      # noqa: E402  (import stubs + reference on sys.path)
from quantum_simulation.mps import MPS  # noqa: E402

store = {}
for tag, (N, d, chi, dt, seed, moves) in {"c128": (8, 2, 6, torch.complex128, 5, (5, 2, 7, 0)),
                                          "f64": (6, 3, 5, torch.float64, 11, (3, 5, 1))}.items():
    m = mg.quiet(MPS, N, d, chi, "cpu", dt, seed)
    for p, t in enumerate(m.B):
        store[f"{tag}/init_B{p}"] = t.detach().numpy()
    for tgt in moves:
        mg.quiet(m.position, tgt)
    for p in range(N):
        store[f"{tag}/A{p}"] = m.A[p].detach().numpy()
        store[f"{tag}/B{p}"] = m.B[p].detach().numpy()
    for p in range(N + 1):
        store[f"{tag}/s{p}"] = m.sWeight[p].detach().numpy()
    store[f"{tag}/cfg"] = np.array([N, d, chi, seed])
    store[f"{tag}/moves"] = np.array(moves)
np.savez_compressed(os.path.join(HERE, "mps.npz"), **store)
print("wrote mps.npz")

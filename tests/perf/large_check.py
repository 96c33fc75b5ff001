#!/usr/bin/env python3
"""BASELINE configs 4 / 5 at (or near) full size on ONE B200: J1-J2 cylinder MPO (W ~ 30-38) at chi=4096 float64 and
Hubbard (d=4, W=6) at chi=2048 complex128.  Checks 64-bit indexing and the large-W mix against the oracle's
tensordot chain executed by cuBLAS on the same device, and reports the matvec rate.  One JSON line per case."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import quantum_simulation_b200 as qsb           # noqa: E402
from quantum_simulation_b200 import _lib        # noqa: E402
from oracle import dmrg_oracle as orc           # noqa: E402

DEV = torch.device("cuda")
cases = [("C4 J1-J2 cylinder 8x6", "j1j2_f64_8x6", torch.float64, int(os.environ.get("C4_CHI", 4096)), 20),
         ("C5 Hubbard JW d=4", "hubbard_f64_L64", torch.complex128, int(os.environ.get("C5_CHI", 2048)), 30)]
for name, key, dtype, chi, p in cases:
    mpo = bench.load_mpo(key, dtype)
    M1, M2 = mpo[p].to(DEV), mpo[p + 1].to(DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    L = torch.randn(M1.shape[0], chi, chi, dtype=dtype, device=DEV, generator=g) / chi ** 0.5
    R = torch.randn(M2.shape[1], chi, chi, dtype=dtype, device=DEV, generator=g) / chi ** 0.5
    x = torch.randn(chi * M1.shape[3] * M2.shape[3] * chi, dtype=dtype, device=DEV, generator=g)
    x /= torch.linalg.norm(x)
    op = qsb.ApplyMPO()
    out = op(x, L, M1, M2, R)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        op(x, L, M1, M2, R, out=out)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    flops = _lib.heff_flops(x, L, M1, M2, R)
    nnz1 = int((M1.abs().sum(dim=(2, 3)) != 0).sum())
    row = {"case": name, "chi": chi, "W": [M1.shape[0], M1.shape[1], M2.shape[1]], "d": M1.shape[2],
           "dtype": str(dtype), "n": x.numel(), "nonzero_blocks_M1": f"{nnz1}/{M1.shape[0] * M1.shape[1]}",
           "ms": ms, "tflops_dense_count": flops / ms / 1e9, "workspace_GB": _lib._workspaces[DEV if DEV.index is not None else torch.device('cuda', torch.cuda.current_device())].numel() / 1e9
           if False else None}
    try:
        ref = orc.apply_mpo(x, L, M1, M2, R)
        row["rel_err_vs_oracle_on_device"] = float(torch.linalg.norm(out - ref) / torch.linalg.norm(ref))
        del ref
    except Exception as ex:      # the oracle's intermediates may not fit next to ours
        row["rel_err_vs_oracle_on_device"] = repr(ex)[:120]
    print(json.dumps(row), flush=True)
    del L, R, x, out
    _lib.release_workspaces()
    torch.cuda.empty_cache()

#!/usr/bin/env python3
"""Time-boxed experiment (NOT part of the product path): FP64 GEMM emulated on the INT8 tensor
pipe (Ozaki-style error-free slicing) at the shape of the H_eff step-1 GEMM of BASELINE config 3.

  C = A B,  A (2048 x 2048) one channel of the left environment,  B (2048 x 8192) theta.
  Every row of A / column of B is scaled by a power of two and cut into s signed 7-bit slices (exact in FP64);
  the s(s+1)/2 slice products with i + j < s are exact INT8 x INT8 -> INT32 GEMMs (K * 127^2 < 2^31), recombined in
  FP64 with their weights 2^(-7(i+j+2)).

What is measured: (1) the relative error of the emulated product against the FP64 product as a function of s -- which s
meets the 1e-12 matvec bar; (2) the time of the s(s+1)/2 INT8 GEMMs alone at LIBRARY speed (torch._int_mm -> cuBLASLt,
the best a hand-written tcgen05 kind::i8 kernel could hope to match), of the slicing and of the recombination -- the
break-even against the DMMA kernel (7.77 ms for 4 channels = 1.94 ms per channel at this shape).

    python tools/ozaki_experiment.py            (one GPU; prints JSON lines)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda"
BITS = 7


def slices_rows(A, s):
    """Row-wise: A[m, :] = 2^e[m] * sum_i q_i[m, :] 2^(-BITS (i + 1)), q_i int8 in [-127, 127]; exact."""
    mx = A.abs().amax(dim=1, keepdim=True).clamp_min(torch.finfo(A.dtype).tiny)
    e = torch.ceil(torch.log2(mx)) + 1.0                   # |A / 2^e| < 1/2 ... strictly below 1
    r = A / torch.exp2(e)
    out = []
    for _ in range(s):
        r = r * float(1 << BITS)
        q = torch.trunc(r)
        r = r - q
        out.append(q.to(torch.int8))
    return out, e.squeeze(1)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    M, K, N = 2048, 2048, 8192
    g = torch.Generator(device=DEV).manual_seed(3)
    cases = {}
    # (a) what bench.py feeds the matvec: dense standard-normal operands
    cases["normal"] = (torch.randn(M, K, dtype=torch.float64, device=DEV, generator=g),
                       torch.randn(K, N, dtype=torch.float64, device=DEV, generator=g))
    # (b) DMRG-like dynamic range: entries spanning 12 decades inside every row / column
    sa = torch.exp(-torch.rand(M, K, dtype=torch.float64, device=DEV, generator=g) * 27.6)
    sb = torch.exp(-torch.rand(K, N, dtype=torch.float64, device=DEV, generator=g) * 27.6)
    cases["wide_range"] = (cases["normal"][0] * sa, cases["normal"][1] * sb)
    t_fp64 = timeit(lambda: cases["normal"][0] @ cases["normal"][1])
    print(json.dumps({"shape": [M, K, N], "cublas_fp64_ms": t_fp64, "dmma_kernel_ms_per_channel_measured_elsewhere": 7.77 / 4}), flush=True)
    for name, (A, B) in cases.items():
        ref = A @ B
        absprod = A.abs() @ B.abs()
        for s in (4, 5, 6, 7, 8):
            qa, ea = slices_rows(A, s)
            qbT, eb = slices_rows(B.T.contiguous(), s)
            qb = [q.T.contiguous() for q in qbT]            # (K, N) int8
            pairs = [(i, j) for i in range(s) for j in range(s) if i + j < s]

            def gemms():
                return [torch._int_mm(qa[i], qb[j]) for (i, j) in pairs]
            prods = gemms()
            C = torch.zeros(M, N, dtype=torch.float64, device=DEV)
            for lvl in range(s - 1, -1, -1):                # smallest weights first
                acc = torch.zeros(M, N, dtype=torch.int64, device=DEV)
                for (i, j), p in zip(pairs, prods):
                    if i + j == lvl:
                        acc += p
                C += acc.to(torch.float64) * 2.0 ** (-BITS * (lvl + 2))
            C = C * torch.exp2(ea).unsqueeze(1) * torch.exp2(eb).unsqueeze(0)
            err = float(torch.linalg.norm(C - ref) / torch.linalg.norm(ref))
            err_abs = float(((C - ref).abs() / absprod).max())
            del prods, C
            t_g = timeit(gemms, 3)
            t_one = timeit(lambda: torch._int_mm(qa[0], qb[0]), 10)
            t_sl = timeit(lambda: slices_rows(B.T.contiguous(), s), 3)
            print(json.dumps({"operands": name, "slices": s, "int8_gemms": len(pairs), "rel_err_fro": err,
                              "max_err_over_absA_absB": err_abs, "int8_gemms_ms": t_g, "one_int8_gemm_ms": t_one,
                              "one_int8_gemm_tops": 2.0 * M * K * N / (t_one * 1e-3) / 1e12,
                              "slice_theta_ms": t_sl, "speedup_vs_dmma_gemms_only": (7.77 / 4) / t_g}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Experiment (not part of the product path): cuSOLVER alternatives for the two-site split at 4096^2 float64 --
syevd (what torch.linalg.eigh calls), syevdx with an index range (upper half of the spectrum only) and gesvdp
(polar-decomposition SVD), called through ctypes on the libcusolver torch already loaded."""
import ctypes as C
import glob
import json
import os
import sys
import time

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
torch.empty(1, device=dev)
paths = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cusolver", "lib", "libcusolver.so*"))
lib = C.CDLL(paths[0] if paths else "libcusolver.so.11")
h = C.c_void_p()
assert lib.cusolverDnCreate(C.byref(h)) == 0
lib.cusolverDnSetStream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
VEC, LOWER, ALL, IDX = 1, 0, 1001, 1002
vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


g = torch.Generator(device=dev).manual_seed(0)
theta = torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)
G = theta @ theta.T
res = {"n": n}
res["torch_eigh_ms"] = timeit(lambda: torch.linalg.eigh(G))
res["gram_gemm_ms"] = timeit(lambda: theta @ theta.T)

W = torch.empty(n, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
lwork = i32(0)
A = G.clone()
lib.cusolverDnDsyevd_bufferSize(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), C.byref(lwork))
work = torch.empty(lwork.value, dtype=torch.float64, device=dev)


def syevd():
    A.copy_(G)
    rc = lib.cusolverDnDsyevd(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), vp(work.data_ptr()), lwork, vp(info.data_ptr()))
    assert rc == 0


res["syevd_ms"] = timeit(syevd)
ref_w = W.clone()

meig = i32(0)
lw2 = i32(0)
il, iu = n // 2 + 1, n
lib.cusolverDnDsyevdx_bufferSize(h, VEC, IDX, LOWER, n, vp(A.data_ptr()), n, dbl(0), dbl(0), il, iu, C.byref(meig),
                                 vp(W.data_ptr()), C.byref(lw2))
work2 = torch.empty(max(lw2.value, 1), dtype=torch.float64, device=dev)


def syevdx():
    A.copy_(G)
    rc = lib.cusolverDnDsyevdx(h, VEC, IDX, LOWER, n, vp(A.data_ptr()), n, dbl(0), dbl(0), il, iu, C.byref(meig),
                               vp(W.data_ptr()), vp(work2.data_ptr()), lw2, vp(info.data_ptr()))
    assert rc == 0


try:
    res["syevdx_upper_half_ms"] = timeit(syevdx)
    res["syevdx_meig"] = meig.value
    res["syevdx_eigs_match"] = float((W[: n // 2] - ref_w[n // 2:]).abs().max() / ref_w[-1])
except Exception as ex:
    res["syevdx_error"] = repr(ex)[:100]

# gesvdp (64-bit API)
try:
    params = vp()
    lib.cusolverDnCreateParams(C.byref(params))
    R64 = 1
    S = torch.empty(n, dtype=torch.float64, device=dev)
    U = torch.empty(n, n, dtype=torch.float64, device=dev)
    V = torch.empty(n, n, dtype=torch.float64, device=dev)
    At = theta.clone()
    wd, wh = C.c_size_t(0), C.c_size_t(0)
    rc = lib.cusolverDnXgesvdp_bufferSize(h, params, VEC, 1, i64(n), i64(n), R64, vp(At.data_ptr()), i64(n), R64,
                                          vp(S.data_ptr()), R64, vp(U.data_ptr()), i64(n), R64, vp(V.data_ptr()), i64(n),
                                          R64, C.byref(wd), C.byref(wh))
    assert rc == 0, rc
    bufd = torch.empty(max(wd.value, 1), dtype=torch.uint8, device=dev)
    bufh = (C.c_char * max(wh.value, 1))()
    herr = dbl(0)

    def gesvdp():
        At.copy_(theta)
        rc = lib.cusolverDnXgesvdp(h, params, VEC, 1, i64(n), i64(n), R64, vp(At.data_ptr()), i64(n), R64, vp(S.data_ptr()),
                                   R64, vp(U.data_ptr()), i64(n), R64, vp(V.data_ptr()), i64(n), R64, vp(bufd.data_ptr()),
                                   wd, bufh, wh, vp(info.data_ptr()), C.byref(herr))
        assert rc == 0, rc
    res["gesvdp_ms"] = timeit(gesvdp)
    Sref = torch.linalg.svdvals(theta)
    res["gesvdp_sv_rel_err"] = float((S - Sref).abs().max() / Sref[0])
    res["gesvdp_h_err_sigma"] = herr.value
except Exception as ex:
    res["gesvdp_error"] = repr(ex)[:160]

from quantum_simulation_b200.svd import two_site_split   # noqa: E402
res["two_site_split_ms"] = timeit(lambda: two_site_split(theta, n // 2, 0.0))
res["qr_n_by_half_ms"] = timeit(lambda: torch.linalg.qr(theta[:, : n // 2]))
print(json.dumps(res))

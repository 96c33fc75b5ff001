// Shared device/host helpers for the qsb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>      // CUtensorMap + enums only; the encode entry point is fetched at run time
#include <stdint.h>
#include <stddef.h>
#include "../../include/qsb200.h"

namespace qsb {

// ---- launch accounting / error capture (host) ------------------------------
extern uint64_t g_count_tma, g_count_cpasync, g_count_other;
extern int g_last_cuda_error;

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return QSB_ERR_LAUNCH; }
    return QSB_OK;
}
inline int check_cuda(cudaError_t e) {
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; (void)cudaGetLastError(); return QSB_ERR_LAUNCH; }
    return QSB_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// the plain batched GEMM every contraction is built from (gemm.cu)
struct GemmArgs {
    int cplx;                       // 0: float64, 1: complex128 (interleaved)
    int M, N, K, P, Z;              // C[z][m][n] = sum_p sum_k A[z,p][m][k] * B[z,p][k][n]
    const void* A; long long a_rs, a_ks, a_ps, a_zs; int conjA;   // element strides; a_rs: stride of m
    const void* B; long long b_rs, b_ks, b_ps, b_zs; int conjB;   // b_rs: stride of n
    void*       C; long long c_ms, c_ns, c_zs;
    int force_path;                 // 0 auto, 1 TMA, 2 cp.async
    // optional k-chunk gating (TMA path only): see TArgs in gemm_tma.cu
    int ksplit = 1, krot = 0, kepoch = 0, klocal = 0;
    int nsplit = 1;                 // > 1: gate whole tiles on the arrival of their COLUMN block of B
    const int* kflags = nullptr;
    int accumulate = 0;             // 1: C += A.B (C must hold valid data), 0: C = A.B
    // optional real scale factors applied in the epilogue: C[z][m][n] = row_scale[m / row_div] * col_scale[n % col_mod] * (A.B)
    // (the Schmidt weights of the theta build, dmrg.py:633-660: rows (l, s) scale with l, columns (t, r) with r)
    const double* row_scale = nullptr; int row_div = 1;
    const double* col_scale = nullptr; int col_mod = 1;
};
int gemm_launch(const GemmArgs& g, cudaStream_t stream);

// ---- small PTX wrappers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor core (SASS: DMMA.8x8x4)
// fragment ownership (lane = 4*g + t):  a = A[g][t], b = B[t][g], d0/d1 = D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src, bool ok) {
    int sz = ok ? BYTES : 0;     // src-size 0 => destination zero-filled, nothing read
    if constexpr (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// ---- mbarrier + TMA ----------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// tiled TMA loads, global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

// ---- tile rasterisation -------------------------------------------------------
// Linear CTA index -> (tile_m, tile_n, batch z).  Tiles are walked in groups of kRasterGroup rows of tiles, n
// fastest inside a group, so that one wave of 148 CTAs covers a ~12 x 12 patch: the A and B panels a wave streams
// from HBM are ~2 * sqrt(148) instead of tiles_m + 148 / tiles_m (L2-friendly rasterisation).
constexpr int kRasterGroup = 12;
__device__ __forceinline__ void tile_coords(int bid, int tiles_m, int tiles_n, int& tm, int& tn, int& z,
                                            int group = kRasterGroup) {
    const int per_z = tiles_m * tiles_n;
    z = bid / per_z;
    const int r = bid - z * per_z;
    const int per_group = group * tiles_n;
    const int grp = r / per_group;
    const int first = grp * group;
    const int gsz = min(group, tiles_m - first);
    const int rr = r - grp * per_group;
    tn = rr / gsz;
    tm = first + (rr - tn * gsz);
}

// ---- reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace qsb

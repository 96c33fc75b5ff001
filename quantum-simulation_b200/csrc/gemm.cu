// FP64 tensor-core (DMMA.8x8x4) batched GEMM core of the qsb200 library.
//
//   C[z][m][n] = sum_p sum_k opA(A[z,p][m][k]) * opB(B[z,p][k][n])
//
// Every dense contraction of the DMRG hot path is one call of this core:
//   H_eff step 1 / step 4   (reference: ncon_torch.py:138 reached from dmrg.py:218-221)
//   env-update step 1 / 3   (reference: ncon_torch.py:138 reached from dmrg.py:275-278, 305-308)
// The reference runs them as torch.tensordot = permute-copy + cuBLAS/MKL GEMM; here the operand
// strides are taken as they are (either dimension of A and of B may be the contiguous one), so no
// permute copies exist.  complex128 is computed from real DMMAs on (re, im) fragments (4M form).
//
// Staging path "cp.async" (this file, kernel gemm_cpasync): multi-stage LDGSTS ring, padded
// shared-memory tiles (conflict-free fragment loads), predicated zero-filled tails, so any extent and
// any 8-byte alignment works.  The TMA-staged kernel for aligned operands lives in gemm_tma.cu.
#include "common.cuh"

namespace qsb {

// ---------------------------------------------------------------------------------------------
// tile configuration
// ---------------------------------------------------------------------------------------------
template <bool CPLX> struct Cfg;
template <> struct Cfg<false> {
    static constexpr int ES = 8;                 // bytes per element
    static constexpr int BM = 128, BN = 128, BK = 16;
    static constexpr int WARPS_M = 2, WARPS_N = 4;
    static constexpr int WM = 64, WN = 32;       // warp tile
    static constexpr int PAD_KC = 4, PAD_RC = 4; // row padding (elements) of the two smem layouts
};
template <> struct Cfg<true> {
    static constexpr int ES = 16;
    static constexpr int BM = 128, BN = 64, BK = 8;
    static constexpr int WARPS_M = 4, WARPS_N = 2;
    static constexpr int WM = 32, WN = 32;
    static constexpr int PAD_KC = 4, PAD_RC = 2;
};
constexpr int kThreads = 256;
constexpr int kStages = 4;

template <bool CPLX, bool KC, int R> struct TileLayout {
    using C = Cfg<CPLX>;
    // KC: [R][BK + pad] (k contiguous);  !KC: [BK][R + pad] (row index contiguous)
    static constexpr int LD = KC ? (C::BK + C::PAD_KC) : (R + C::PAD_RC);
    static constexpr int ELEMS = KC ? R * LD : C::BK * LD;
    static constexpr int BYTES = ELEMS * C::ES;
};
template <bool CPLX, int R> constexpr int tile_bytes_max() {
    return TileLayout<CPLX, true, R>::BYTES > TileLayout<CPLX, false, R>::BYTES
               ? TileLayout<CPLX, true, R>::BYTES : TileLayout<CPLX, false, R>::BYTES;
}
template <bool CPLX> constexpr int stage_bytes() {
    return tile_bytes_max<CPLX, Cfg<CPLX>::BM>() + tile_bytes_max<CPLX, Cfg<CPLX>::BN>();
}

struct KArgs {
    int M, N, K, P;
    int tiles_m, tiles_n;
    const char* A; long long a_rs, a_ks, a_ps, a_zs;
    const char* B; long long b_rs, b_ks, b_ps, b_zs;
    char* C; long long c_ms, c_ns, c_zs;
    double sgnA, sgnB;              // -1 when the operand is conjugated (complex only)
    int accum;                      // 1: C += A.B
    const double* row_scale; int row_div;    // epilogue scale factors (nullptr: none)
    const double* col_scale; int col_mod;
};

// one operand tile: global -> shared with cp.async, zero fill outside [r_valid) x [k_valid)
template <bool CPLX, int VEC, bool KC, int R>
__device__ __forceinline__ void load_tile(uint32_t sdst, const char* gorigin, const char* gsafe,
                                          long long rs, long long ks, int r_valid, int k_valid, int tid) {
    using C = Cfg<CPLX>;
    using TL = TileLayout<CPLX, KC, R>;
    constexpr int EPV = VEC / C::ES;             // elements per cp.async
    if constexpr (KC) {
        constexpr int CPR = C::BK / EPV;         // copies per tile row
        constexpr int TOTAL = R * CPR;
        static_assert(TOTAL % kThreads == 0, "tile copy count");
#pragma unroll
        for (int i = 0; i < TOTAL / kThreads; ++i) {
            const int c = tid + i * kThreads;
            const int r = c / CPR, k = (c % CPR) * EPV;
            const bool ok = (r < r_valid) && (k < k_valid);
            const char* src = ok ? gorigin + ((long long)r * rs + k) * C::ES : gsafe;
            cp_async<VEC>(sdst + (uint32_t)(r * TL::LD + k) * C::ES, src, ok);
        }
    } else {
        constexpr int CPK = R / EPV;             // copies per k row
        constexpr int TOTAL = C::BK * CPK;
        static_assert(TOTAL % kThreads == 0, "tile copy count");
#pragma unroll
        for (int i = 0; i < TOTAL / kThreads; ++i) {
            const int c = tid + i * kThreads;
            const int k = c / CPK, r = (c % CPK) * EPV;
            const bool ok = (r < r_valid) && (k < k_valid);
            const char* src = ok ? gorigin + ((long long)k * ks + r) * C::ES : gsafe;
            cp_async<VEC>(sdst + (uint32_t)(k * TL::LD + r) * C::ES, src, ok);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool CPLX, bool A_KC, bool B_KC, int VEC>
__global__ void __launch_bounds__(kThreads, 1) gemm_cpasync(const KArgs g) {
    using C = Cfg<CPLX>;
    using TA = TileLayout<CPLX, A_KC, C::BM>;
    using TB = TileLayout<CPLX, B_KC, C::BN>;
    constexpr int A_BYTES = tile_bytes_max<CPLX, C::BM>();
    constexpr int STAGE = stage_bytes<CPLX>();
    constexpr int MI = C::WM / 8, NJ = C::WN / 8;
    constexpr int ACC = CPLX ? 4 : 2;

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp / C::WARPS_N) * C::WM;
    const int wn0 = (warp % C::WARPS_N) * C::WN;

    int tm, tn, z;
    tile_coords(blockIdx.x, g.tiles_m, g.tiles_n, tm, tn, z);
    const int m0 = tm * C::BM, n0 = tn * C::BN;

    const char* Az = g.A + (g.a_zs * z + g.a_rs * m0) * C::ES;
    const char* Bz = g.B + (g.b_zs * z + g.b_rs * n0) * C::ES;
    const int m_valid = g.M - m0, n_valid = g.N - n0;

    const int ktiles = (g.K + C::BK - 1) / C::BK;
    const int iters = ktiles * g.P;

    auto issue = [&](int it) {
        if (it < iters) {
            const int p = it / ktiles, k0 = (it - p * ktiles) * C::BK;
            const uint32_t sa = sbase + (it % kStages) * STAGE, sb = sa + A_BYTES;
            load_tile<CPLX, VEC, A_KC, C::BM>(sa, Az + (g.a_ps * p + g.a_ks * k0) * C::ES, g.A,
                                              g.a_rs, g.a_ks, m_valid, g.K - k0, tid);
            load_tile<CPLX, VEC, B_KC, C::BN>(sb, Bz + (g.b_ps * p + g.b_ks * k0) * C::ES, g.B,
                                              g.b_rs, g.b_ks, n_valid, g.K - k0, tid);
        }
        cp_async_commit();
    };

    double acc[MI][NJ][ACC];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int q = 0; q < ACC; ++q) acc[i][j][q] = 0.0;

    // per-thread fragment offsets (bytes) inside a stage
    const uint32_t a_off = A_KC ? (uint32_t)((wm0 + g8) * TA::LD + t4) * C::ES
                                : (uint32_t)(t4 * TA::LD + wm0 + g8) * C::ES;
    const uint32_t b_off = A_BYTES + (B_KC ? (uint32_t)((wn0 + g8) * TB::LD + t4) * C::ES
                                           : (uint32_t)(t4 * TB::LD + wn0 + g8) * C::ES);
    constexpr uint32_t A_I = (A_KC ? 8 * TA::LD : 8) * C::ES;       // next 8 rows of the warp tile
    constexpr uint32_t A_K = (A_KC ? 4 : 4 * TA::LD) * C::ES;       // next k4 step
    constexpr uint32_t B_J = (B_KC ? 8 * TB::LD : 8) * C::ES;
    constexpr uint32_t B_K = (B_KC ? 4 : 4 * TB::LD) * C::ES;

#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) issue(s);

    for (int it = 0; it < iters; ++it) {
        cp_async_wait<kStages - 2>();
        __syncthreads();                       // stage `it` landed for everyone; stage it-1 is free
        issue(it + kStages - 1);
        const unsigned char* st = smem + (it % kStages) * STAGE;
        constexpr int KS = C::BK / 4;
        if constexpr (!CPLX) {
            double a[2][MI], b[2][NJ];       // register double buffer over the k4 steps
#pragma unroll
            for (int i = 0; i < MI; ++i) a[0][i] = *reinterpret_cast<const double*>(st + a_off + i * A_I);
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[0][j] = *reinterpret_cast<const double*>(st + b_off + j * B_J);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < KS) {
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        a[nxt][i] = *reinterpret_cast<const double*>(st + a_off + i * A_I + (kk + 1) * A_K);
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        b[nxt][j] = *reinterpret_cast<const double*>(st + b_off + j * B_J + (kk + 1) * B_K);
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
            }
        } else {
            double2 a[2][MI], b[2][NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[0][i] = *reinterpret_cast<const double2*>(st + a_off + i * A_I);
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[0][j] = *reinterpret_cast<const double2*>(st + b_off + j * B_J);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < KS) {
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        a[nxt][i] = *reinterpret_cast<const double2*>(st + a_off + i * A_I + (kk + 1) * A_K);
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        b[nxt][j] = *reinterpret_cast<const double2*>(st + b_off + j * B_J + (kk + 1) * B_K);
                }
                double ai[MI], nai[MI], bi[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) { ai[i] = a[cur][i].y * g.sgnA; nai[i] = -ai[i]; }
#pragma unroll
                for (int j = 0; j < NJ; ++j) bi[j] = b[cur][j].y * g.sgnB;
                // re += ar*br ; im += ar*bi   then   re += (-ai)*bi ; im += ai*br
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884(acc[i][j][0], acc[i][j][1], a[cur][i].x, b[cur][j].x);
                        dmma884(acc[i][j][2], acc[i][j][3], a[cur][i].x, bi[j]);
                    }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884(acc[i][j][0], acc[i][j][1], nai[i], bi[j]);
                        dmma884(acc[i][j][2], acc[i][j][3], ai[i], b[cur][j].x);
                    }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row = g8 + 8i][col = 2*t4 + 8j + {0,1}] of its warp tile
    char* Cz = g.C + g.c_zs * z * C::ES;
    const bool vec_ok = (g.c_ns == 1) && ((g.c_ms & 1) == 0 || CPLX) && ((g.c_zs & 1) == 0 || CPLX) &&
                        ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + wm0 + 8 * i + g8;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = n0 + wn0 + 8 * j + 2 * t4;
            if (n >= g.N) continue;
            char* dst = Cz + (g.c_ms * m + g.c_ns * n) * C::ES;
            if (g.row_scale || g.col_scale) {
                const double rs = g.row_scale ? g.row_scale[m / g.row_div] : 1.0;
                const double s0 = rs * (g.col_scale ? g.col_scale[n % g.col_mod] : 1.0);
                const double s1 = rs * ((g.col_scale && n + 1 < g.N) ? g.col_scale[(n + 1) % g.col_mod] : 1.0);
                acc[i][j][0] *= s0; acc[i][j][1] *= s1;
                if constexpr (CPLX) { acc[i][j][2] *= s0; acc[i][j][3] *= s1; }
            }
            if constexpr (!CPLX) {
                if (g.accum) {
                    acc[i][j][0] += *reinterpret_cast<const double*>(dst);
                    if (n + 1 < g.N) acc[i][j][1] += *reinterpret_cast<const double*>(dst + g.c_ns * C::ES);
                }
                if (vec_ok && n + 1 < g.N) {
                    *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    *reinterpret_cast<double*>(dst) = acc[i][j][0];
                    if (n + 1 < g.N) *reinterpret_cast<double*>(dst + g.c_ns * C::ES) = acc[i][j][1];
                }
            } else {
                if (g.accum) {
                    const double2 o0 = *reinterpret_cast<const double2*>(dst);
                    acc[i][j][0] += o0.x; acc[i][j][2] += o0.y;
                    if (n + 1 < g.N) {
                        const double2 o1 = *reinterpret_cast<const double2*>(dst + g.c_ns * C::ES);
                        acc[i][j][1] += o1.x; acc[i][j][3] += o1.y;
                    }
                }
                *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][2]);
                if (n + 1 < g.N)
                    *reinterpret_cast<double2*>(dst + g.c_ns * C::ES) = make_double2(acc[i][j][1], acc[i][j][3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
template <bool CPLX, bool A_KC, bool B_KC, int VEC>
static int launch_cpasync(const KArgs& k, int Z, cudaStream_t stream) {
    auto kern = gemm_cpasync<CPLX, A_KC, B_KC, VEC>;
    constexpr int SMEM = kStages * stage_bytes<CPLX>();
    static bool configured[64] = {};         // per instantiation and device (the attribute is per device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return QSB_ERR_DEVICE;
    if (!configured[dev]) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        if (rc) return rc;
        configured[dev] = true;
    }
    const long long blocks = (long long)k.tiles_m * k.tiles_n * Z;
    if (blocks <= 0) return QSB_OK;
    if (blocks > 2147483647LL) return QSB_ERR_SHAPE;
    kern<<<(unsigned)blocks, kThreads, SMEM, stream>>>(k);
    ++g_count_cpasync;
    return check_launch();
}

int gemm_tma_try(const GemmArgs& g, cudaStream_t stream, bool* used);   // gemm_tma.cu

int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
    if (g.M < 0 || g.N < 0 || g.K < 0 || g.P < 0 || g.Z < 0) return QSB_ERR_SHAPE;
    if (g.M == 0 || g.N == 0 || g.Z == 0) return QSB_OK;
    if (!g.A || !g.B || !g.C) return QSB_ERR_ARG;
    const bool cplx = g.cplx != 0;

    if (g.force_path != 2) {
        bool used = false;
        int rc = gemm_tma_try(g, stream, &used);
        if (rc) return rc;
        if (used) return QSB_OK;
        if (g.force_path == 1) return QSB_ERR_STRIDE;
    }
    if (g.ksplit > 1 || g.nsplit > 1) return QSB_ERR_STRIDE;     // chunk gating needs the TMA-staged kernel

    KArgs k;
    k.M = g.M; k.N = g.N; k.K = g.K; k.P = g.P;
    k.A = (const char*)g.A; k.a_rs = g.a_rs; k.a_ks = g.a_ks; k.a_ps = g.a_ps; k.a_zs = g.a_zs;
    k.B = (const char*)g.B; k.b_rs = g.b_rs; k.b_ks = g.b_ks; k.b_ps = g.b_ps; k.b_zs = g.b_zs;
    k.C = (char*)g.C; k.c_ms = g.c_ms; k.c_ns = g.c_ns; k.c_zs = g.c_zs;
    k.sgnA = g.conjA ? -1.0 : 1.0; k.sgnB = g.conjB ? -1.0 : 1.0;
    k.accum = g.accumulate;
    k.row_scale = g.row_scale; k.row_div = g.row_div > 0 ? g.row_div : 1;
    k.col_scale = g.col_scale; k.col_mod = g.col_mod > 0 ? g.col_mod : 1;
    // extent-1 dimensions carry arbitrary strides in torch: normalise them
    if (g.K == 1) { k.a_ks = 1; k.b_ks = 1; }
    if (g.M == 1 && k.a_ks != 1) k.a_rs = 1;
    if (g.N == 1 && k.b_ks != 1) k.b_rs = 1;
    if (g.P <= 1) { k.a_ps = 0; k.b_ps = 0; }
    if (g.Z <= 1) { k.a_zs = 0; k.b_zs = 0; k.c_zs = 0; }
    const bool a_kc = (k.a_ks == 1), b_kc = (k.b_ks == 1);
    if (!a_kc && k.a_rs != 1) return QSB_ERR_STRIDE;
    if (!b_kc && k.b_rs != 1) return QSB_ERR_STRIDE;
    const int BM = cplx ? Cfg<true>::BM : Cfg<false>::BM, BN = cplx ? Cfg<true>::BN : Cfg<false>::BN;
    k.tiles_m = (g.M + BM - 1) / BM;
    k.tiles_n = (g.N + BN - 1) / BN;

    if (cplx) {
        if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C)) return QSB_ERR_STRIDE;
        if (a_kc) return b_kc ? launch_cpasync<true, true, true, 16>(k, g.Z, stream)
                              : launch_cpasync<true, true, false, 16>(k, g.Z, stream);
        return b_kc ? launch_cpasync<true, false, true, 16>(k, g.Z, stream)
                    : launch_cpasync<true, false, false, 16>(k, g.Z, stream);
    }
    // float64: 16-byte copies need even strides/extents along every non-contiguous step
    auto even = [](long long v) { return (v & 1) == 0; };
    bool v16 = aligned16(g.A) && aligned16(g.B);
    v16 = v16 && even(a_kc ? g.K : g.M) && even(a_kc ? k.a_rs : k.a_ks) && even(k.a_ps) && even(k.a_zs);
    v16 = v16 && even(b_kc ? g.K : g.N) && even(b_kc ? k.b_rs : k.b_ks) && even(k.b_ps) && even(k.b_zs);
    if (v16) {
        if (a_kc) return b_kc ? launch_cpasync<false, true, true, 16>(k, g.Z, stream)
                              : launch_cpasync<false, true, false, 16>(k, g.Z, stream);
        return b_kc ? launch_cpasync<false, false, true, 16>(k, g.Z, stream)
                    : launch_cpasync<false, false, false, 16>(k, g.Z, stream);
    }
    if (a_kc) return b_kc ? launch_cpasync<false, true, true, 8>(k, g.Z, stream)
                          : launch_cpasync<false, true, false, 8>(k, g.Z, stream);
    return b_kc ? launch_cpasync<false, false, true, 8>(k, g.Z, stream)
                : launch_cpasync<false, false, false, 8>(k, g.Z, stream);
}

// ---------------------------------------------------------------------------------------------
// DMMA issue-rate probe: register-resident mma.sync.m8n8k4.f64 on every SM
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_dmma_kernel(int iters, double* sink) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;           // never true; keeps the chain alive
}

}  // namespace qsb

using namespace qsb;

extern "C" int qsb_gemm(const qsb_gemm_desc* d, void* stream) {
    if (!d) return QSB_ERR_ARG;
    if (d->dtype != QSB_F64 && d->dtype != QSB_C128) return QSB_ERR_DTYPE;
    GemmArgs g;
    g.cplx = d->dtype == QSB_C128;
    g.M = d->M; g.N = d->N; g.K = d->K; g.P = d->P; g.Z = d->Z;
    g.A = d->A; g.a_rs = d->a_ms; g.a_ks = d->a_ks; g.a_ps = d->a_ps; g.a_zs = d->a_zs; g.conjA = d->conjA;
    g.B = d->B; g.b_rs = d->b_ns; g.b_ks = d->b_ks; g.b_ps = d->b_ps; g.b_zs = d->b_zs; g.conjB = d->conjB;
    g.C = d->C; g.c_ms = d->c_ms; g.c_ns = d->c_ns; g.c_zs = d->c_zs;
    g.force_path = d->force_path;
    g.accumulate = d->accumulate;
    g.row_scale = d->row_scale; g.row_div = d->row_div; g.col_scale = d->col_scale; g.col_mod = d->col_mod;
    if ((g.row_scale && g.row_div < 1) || (g.col_scale && g.col_mod < 1)) return QSB_ERR_ARG;
    if ((g.row_scale || g.col_scale) && g.accumulate) return QSB_ERR_ARG;
    if (g.P < 1) g.P = 1;
    return gemm_launch(g, (cudaStream_t)stream);
}

extern "C" int qsb_probe_dmma(int iters, double* flops, void* sink, void* stream) {
    if (iters <= 0 || !sink) return QSB_ERR_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return QSB_ERR_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 2;              // 16 warps per SM, 4 per scheduler
    probe_dmma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, (double*)sink);
    ++g_count_other;
    if (flops) *flops = (double)blocks * 8.0 /*warps*/ * (double)iters * 8.0 /*mma*/ * 512.0;
    return check_launch();
}

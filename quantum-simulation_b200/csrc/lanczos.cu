// Fused Lanczos vector kernels (HBM-bound).
//
// They replace the torch BLAS-1/2 calls of _LanczosEngine.forward, reference
// quantum_simulation/eigensolver/lanczos_solver.py:
//   qsb_lz_dot        vdot :241 (alpha), K^H w of the PRO step :254, linalg.norm :259 / :149
//   qsb_lz_axpy_norm  the two add_ of the three-term recurrence :245-247, the PRO projection
//                     :255-256, fused with the norm :259
//   qsb_lz_scale      copy_ + div_ :268-270 and _normalize_to :148-154 (one pass)
//   qsb_lz_combine    Ritz vector addmv :293 fused with the norm of :294
// Scalars (alpha, beta, overlaps) stay on the device in `scal` ((re, im) double pairs); reductions are
// two-stage with a fixed grid and a fixed summation order (bit-reproducible run to run).
// Loads are 16-byte vectorised and coalesced; partial sums use warp shuffles.
#include "common.cuh"

namespace qsb {

constexpr int kLzThreads = 256;
constexpr int kLzMaxBlocks = 148 * 8;
constexpr int kLzMaxVec = 4;          // vectors handled per pass over w
constexpr int kLzMaxCombine = 64;

// MODE 0: real, one scalar per unit; MODE 1: real, two scalars per unit; MODE 2: complex (re, im)
template <typename T, int MODE> struct Unit {
    double a, b;
    __device__ __forceinline__ static Unit load(const T* p, long long u) {
        Unit r;
        if constexpr (MODE == 0) { r.a = (double)p[u]; r.b = 0.0; }
        else if constexpr (sizeof(T) == 8) { const double2 v = reinterpret_cast<const double2*>(p)[u]; r.a = v.x; r.b = v.y; }
        else { const float2 v = reinterpret_cast<const float2*>(p)[u]; r.a = v.x; r.b = v.y; }
        return r;
    }
    __device__ __forceinline__ void store(T* p, long long u) const {
        if constexpr (MODE == 0) { p[u] = (T)a; }
        else if constexpr (sizeof(T) == 8) { reinterpret_cast<double2*>(p)[u] = make_double2(a, b); }
        else { reinterpret_cast<float2*>(p)[u] = make_float2((float)a, (float)b); }
    }
};

struct Scratch {            // device layout of the caller's scratch buffer
    unsigned int ticket;
    unsigned int pad[63];
    double partial[1];      // [kLzMaxBlocks][kLzMaxVec][2]
};

// block-level sum of `cnt` doubles per thread -> thread 0 of the block holds the totals in v[]
template <int CNT>
__device__ __forceinline__ void block_sum(double (&v)[CNT], double* sm /* [8][CNT] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CNT; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) sm[warp * CNT + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) {
            double s = 0.0;
            for (int w = 0; w < kLzThreads / 32; ++w) s += sm[w * CNT + i];
            v[i] = s;
        }
    }
}

// second stage: the last block to finish sums the per-block partials in block order
template <int CNT, typename F>
__device__ __forceinline__ void grid_finish(double (&v)[CNT], Scratch* sc, double* sm, F&& write_result) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) sc->partial[(size_t)blockIdx.x * CNT + i] = v[i];
        __threadfence();
        const unsigned int t = atomicAdd(&sc->ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double acc[CNT];
#pragma unroll
    for (int i = 0; i < CNT; ++i) acc[i] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kLzThreads) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) acc[i] += __ldcg(&sc->partial[(size_t)b * CNT + i]);
    }
    __syncthreads();
    block_sum<CNT>(acc, sm);
    if (threadIdx.x == 0) {
        write_result(acc);
        sc->ticket = 0;
    }
}

// ---- dot:  scal[out_idx + i] = <X_i, w> ------------------------------------------------------------
template <typename T, int MODE, int MV>
__global__ void __launch_bounds__(kLzThreads) lz_dot_kernel(const T* X, long long ldx_units, const T* w,
                                                            long long nunits, double* scal, int out_idx, int flags,
                                                            Scratch* sc) {
    __shared__ double sm[(kLzThreads / 32) * MV * 2];
    double v[MV * 2];
#pragma unroll
    for (int i = 0; i < MV * 2; ++i) v[i] = 0.0;
    const long long stride = (long long)gridDim.x * kLzThreads;
#pragma unroll 2
    for (long long u = (long long)blockIdx.x * kLzThreads + threadIdx.x; u < nunits; u += stride) {
        const Unit<T, MODE> ww = Unit<T, MODE>::load(w, u);
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            const Unit<T, MODE> x = Unit<T, MODE>::load(X, u + i * ldx_units);
            v[2 * i] += x.a * ww.a + x.b * ww.b;
            if constexpr (MODE == 2) v[2 * i + 1] += x.a * ww.b - x.b * ww.a;   // conj(x) * w
        }
    }
    block_sum<MV * 2>(v, sm);
    __syncthreads();
    grid_finish<MV * 2>(v, sc, sm, [&](double (&acc)[MV * 2]) {
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            double re = acc[2 * i], im = acc[2 * i + 1];
            if (flags & 1) im = 0.0;
            if (flags & 2) { re = sqrt(re); im = 0.0; }
            scal[2 * (out_idx + i)] = re;
            scal[2 * (out_idx + i) + 1] = im;
        }
    });
}

// ---- axpy + norm:  w -= sum_i c_i X_i ; scal[norm_idx] = ||w|| -------------------------------------
template <typename T, int MODE, int MV>
__global__ void __launch_bounds__(kLzThreads) lz_axpy_norm_kernel(const T* X, long long ldx_units, T* w,
                                                                  long long nunits, double* scal, int coef_idx,
                                                                  int norm_idx, int flags, Scratch* sc) {
    __shared__ double sm[(kLzThreads / 32)];
    double cr[MV], ci[MV];
#pragma unroll
    for (int i = 0; i < MV; ++i) { cr[i] = scal[2 * (coef_idx + i)]; ci[i] = scal[2 * (coef_idx + i) + 1]; }
    double v[1] = {0.0};
    const long long stride = (long long)gridDim.x * kLzThreads;
#pragma unroll 2
    for (long long u = (long long)blockIdx.x * kLzThreads + threadIdx.x; u < nunits; u += stride) {
        Unit<T, MODE> ww = Unit<T, MODE>::load(w, u);
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            const Unit<T, MODE> x = Unit<T, MODE>::load(X, u + i * ldx_units);
            if constexpr (MODE == 2) {
                ww.a -= cr[i] * x.a - ci[i] * x.b;
                ww.b -= cr[i] * x.b + ci[i] * x.a;
            } else {
                ww.a -= cr[i] * x.a;
                ww.b -= cr[i] * x.b;
            }
        }
        ww.store(w, u);
        if constexpr (sizeof(T) == 4) ww = Unit<T, MODE>::load(w, u);     // norm of the rounded value
        v[0] += ww.a * ww.a + ww.b * ww.b;
    }
    if (norm_idx < 0) return;
    block_sum<1>(v, sm);
    __syncthreads();
    grid_finish<1>(v, sc, sm, [&](double (&acc)[1]) {
        scal[2 * norm_idx] = (flags & QSB_LZ_NOSQRT) ? acc[0] : sqrt(acc[0]);
        scal[2 * norm_idx + 1] = 0.0;
    });
}

// ---- scale:  dst = src / Re(scal[idx]) --------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kLzThreads) lz_scale_kernel(const T* src, T* dst, long long nunits,
                                                              const double* scal, int idx) {
    const double inv = 1.0 / scal[2 * idx];      // one division; the stream multiplies (<= 1.5 ulp from x / s)
    const long long stride = (long long)gridDim.x * kLzThreads;
#pragma unroll 4
    for (long long u = (long long)blockIdx.x * kLzThreads + threadIdx.x; u < nunits; u += stride) {
        Unit<T, MODE> x = Unit<T, MODE>::load(src, u);
        x.a *= inv; x.b *= inv;
        x.store(dst, u);
    }
}

// ---- combine:  dst = sum_i y_i X_i ; scal[norm_idx] = ||dst|| --------------------------------------
struct CombineCoef { double y[kLzMaxCombine]; };

template <typename T, int MODE>
__global__ void __launch_bounds__(kLzThreads) lz_combine_kernel(const T* X, long long ldx_units, int m,
                                                                const CombineCoef cf, T* dst, long long nunits,
                                                                double* scal, int norm_idx, int flags, Scratch* sc) {
    __shared__ double sm[(kLzThreads / 32)];
    double v[1] = {0.0};
    const long long stride = (long long)gridDim.x * kLzThreads;
    for (long long u = (long long)blockIdx.x * kLzThreads + threadIdx.x; u < nunits; u += stride) {
        Unit<T, MODE> acc; acc.a = 0.0; acc.b = 0.0;
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const Unit<T, MODE> x = Unit<T, MODE>::load(X, u + i * ldx_units);
            acc.a += cf.y[i] * x.a;
            acc.b += cf.y[i] * x.b;
        }
        acc.store(dst, u);
        if constexpr (sizeof(T) == 4) acc = Unit<T, MODE>::load(dst, u);
        v[0] += acc.a * acc.a + acc.b * acc.b;
    }
    if (norm_idx < 0) return;
    block_sum<1>(v, sm);
    __syncthreads();
    grid_finish<1>(v, sc, sm, [&](double (&acc)[1]) {
        scal[2 * norm_idx] = (flags & QSB_LZ_NOSQRT) ? acc[0] : sqrt(acc[0]);
        scal[2 * norm_idx + 1] = 0.0;
    });
}

// ---- host side -----------------------------------------------------------------------------------
struct VecView {
    bool ok; bool f32; int mode; long long nunits; long long ld_mul;   // ldx_units = ldx * ld_mul / ld_div
    int ld_div;
};

// choose the access mode for vectors of n elements; every pointer in `ptrs` and the row pitch must allow it
static VecView vec_view(int dtype, long long n, long long ldx, const void* const* ptrs, int nptr) {
    VecView v{};
    v.ok = true;
    v.f32 = (dtype == QSB_F32 || dtype == QSB_C64);
    const bool cplx = (dtype == QSB_C128 || dtype == QSB_C64);
    const size_t unit2 = v.f32 ? 8 : 16;
    bool al = true;
    for (int i = 0; i < nptr; ++i) al = al && ((reinterpret_cast<uintptr_t>(ptrs[i]) % unit2) == 0);
    if (cplx) {
        if (!al) { v.ok = false; return v; }
        v.mode = 2; v.nunits = n; v.ld_mul = 1; v.ld_div = 1;
    } else if (al && (n % 2 == 0) && (ldx % 2 == 0)) {
        v.mode = 1; v.nunits = n / 2; v.ld_mul = 1; v.ld_div = 2;
    } else {
        v.mode = 0; v.nunits = n; v.ld_mul = 1; v.ld_div = 1;
    }
    return v;
}

static int lz_grid(long long nunits) {
    long long b = (nunits + kLzThreads * 4 - 1) / (kLzThreads * 4);
    if (b < 1) b = 1;
    if (b > kLzMaxBlocks) b = kLzMaxBlocks;
    return (int)b;
}

#define QSB_LZ_DISPATCH(VIEW, ...)                                                    \
    do {                                                                              \
        if ((VIEW).f32) {                                                             \
            using T = float;                                                          \
            if ((VIEW).mode == 0) { constexpr int MODE = 0; __VA_ARGS__; }                   \
            else if ((VIEW).mode == 1) { constexpr int MODE = 1; __VA_ARGS__; }              \
            else { constexpr int MODE = 2; __VA_ARGS__; }                                    \
        } else {                                                                      \
            using T = double;                                                         \
            if ((VIEW).mode == 0) { constexpr int MODE = 0; __VA_ARGS__; }                   \
            else if ((VIEW).mode == 1) { constexpr int MODE = 1; __VA_ARGS__; }              \
            else { constexpr int MODE = 2; __VA_ARGS__; }                                    \
        }                                                                             \
    } while (0)

static bool lz_dtype_ok(int dtype) { return dtype >= QSB_F64 && dtype <= QSB_C64; }
static size_t lz_esize(int dtype) {
    switch (dtype) { case QSB_F64: return 8; case QSB_C128: return 16; case QSB_F32: return 4; default: return 8; }
}

}  // namespace qsb

using namespace qsb;

extern "C" size_t qsb_lz_scratch_bytes(int max_vectors) {
    (void)max_vectors;     // passes are chunked to kLzMaxVec vectors, the scratch size is fixed
    return sizeof(unsigned int) * 64 + sizeof(double) * (size_t)kLzMaxBlocks * kLzMaxVec * 2;
}

extern "C" int qsb_lz_dot(int dtype, int64_t n, int m, const void* X, int64_t ldx, const void* w, double* scal,
                          int out_idx, int flags, void* scratch, void* stream_) {
    if (!lz_dtype_ok(dtype)) return QSB_ERR_DTYPE;
    if (n < 0 || m < 0 || out_idx < 0) return QSB_ERR_ARG;
    if (m == 0) return QSB_OK;
    if (!X || !w || !scal || !scratch) return QSB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const void* ptrs[2] = {X, w};
    const VecView v = vec_view(dtype, n, m > 1 ? ldx : 0, ptrs, 2);
    if (!v.ok) return QSB_ERR_STRIDE;
    const long long ldu = ldx * v.ld_mul / v.ld_div;
    const int grid = lz_grid(v.nunits);
    Scratch* sc = (Scratch*)scratch;
    const size_t es = lz_esize(dtype);
    int i0 = 0;
    while (i0 < m) {
        const int left = m - i0;
        const int mv = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
        const char* Xi = (const char*)X + (size_t)i0 * ldx * es;
        QSB_LZ_DISPATCH(v, {
            if (mv == 4) lz_dot_kernel<T, MODE, 4><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (const T*)w, v.nunits, scal, out_idx + i0, flags, sc);
            else if (mv == 2) lz_dot_kernel<T, MODE, 2><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (const T*)w, v.nunits, scal, out_idx + i0, flags, sc);
            else lz_dot_kernel<T, MODE, 1><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (const T*)w, v.nunits, scal, out_idx + i0, flags, sc);
        });
        ++g_count_other;
        int rc = check_launch();
        if (rc) return rc;
        i0 += mv;
    }
    return QSB_OK;
}

extern "C" int qsb_lz_axpy_norm(int dtype, int64_t n, int m, const void* X, int64_t ldx, void* w, double* scal,
                                int coef_idx, int norm_idx, int flags, void* scratch, void* stream_) {
    if (!lz_dtype_ok(dtype)) return QSB_ERR_DTYPE;
    if (n < 0 || m < 1 || coef_idx < 0) return QSB_ERR_ARG;
    if (!X || !w || !scal || !scratch) return QSB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const void* ptrs[2] = {X, w};
    const VecView v = vec_view(dtype, n, m > 1 ? ldx : 0, ptrs, 2);
    if (!v.ok) return QSB_ERR_STRIDE;
    const long long ldu = ldx * v.ld_mul / v.ld_div;
    const int grid = lz_grid(v.nunits);
    Scratch* sc = (Scratch*)scratch;
    const size_t es = lz_esize(dtype);
    int i0 = 0;
    while (i0 < m) {
        const int left = m - i0;
        const int mv = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
        const int nidx = (i0 + mv >= m) ? norm_idx : -1;    // the norm belongs to the last pass
        const char* Xi = (const char*)X + (size_t)i0 * ldx * es;
        QSB_LZ_DISPATCH(v, {
            if (mv == 4) lz_axpy_norm_kernel<T, MODE, 4><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (T*)w, v.nunits, scal, coef_idx + i0, nidx, flags, sc);
            else if (mv == 2) lz_axpy_norm_kernel<T, MODE, 2><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (T*)w, v.nunits, scal, coef_idx + i0, nidx, flags, sc);
            else lz_axpy_norm_kernel<T, MODE, 1><<<grid, kLzThreads, 0, stream>>>((const T*)Xi, ldu, (T*)w, v.nunits, scal, coef_idx + i0, nidx, flags, sc);
        });
        ++g_count_other;
        int rc = check_launch();
        if (rc) return rc;
        i0 += mv;
    }
    return QSB_OK;
}

extern "C" int qsb_lz_scale(int dtype, int64_t n, const void* src, void* dst, const double* scal, int idx,
                            void* stream_) {
    if (!lz_dtype_ok(dtype)) return QSB_ERR_DTYPE;
    if (n < 0 || idx < 0) return QSB_ERR_ARG;
    if (n == 0) return QSB_OK;
    if (!src || !dst || !scal) return QSB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const void* ptrs[2] = {src, dst};
    const VecView v = vec_view(dtype, n, 0, ptrs, 2);
    if (!v.ok) return QSB_ERR_STRIDE;
    const int grid = lz_grid(v.nunits);
    QSB_LZ_DISPATCH(v, {
        lz_scale_kernel<T, MODE><<<grid, kLzThreads, 0, stream>>>((const T*)src, (T*)dst, v.nunits, scal, idx);
    });
    ++g_count_other;
    return check_launch();
}

extern "C" int qsb_lz_combine(int dtype, int64_t n, int m, const void* X, int64_t ldx, const double* y_host,
                              void* dst, double* scal, int norm_idx, int flags, void* scratch, void* stream_) {
    if (!lz_dtype_ok(dtype)) return QSB_ERR_DTYPE;
    if (n < 0 || m < 1 || m > kLzMaxCombine) return QSB_ERR_ARG;
    if (!X || !y_host || !dst || !scal || !scratch) return QSB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const void* ptrs[2] = {X, dst};
    const VecView v = vec_view(dtype, n, m > 1 ? ldx : 0, ptrs, 2);
    if (!v.ok) return QSB_ERR_STRIDE;
    const long long ldu = ldx * v.ld_mul / v.ld_div;
    const int grid = lz_grid(v.nunits);
    CombineCoef cf;
    for (int i = 0; i < kLzMaxCombine; ++i) cf.y[i] = i < m ? y_host[i] : 0.0;
    Scratch* sc = (Scratch*)scratch;
    QSB_LZ_DISPATCH(v, {
        lz_combine_kernel<T, MODE><<<grid, kLzThreads, 0, stream>>>((const T*)X, ldu, m, cf, (T*)dst, v.nunits, scal, norm_idx, flags, sc);
    });
    ++g_count_other;
    return check_launch();
}

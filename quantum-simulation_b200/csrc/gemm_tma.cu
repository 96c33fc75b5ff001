// TMA-staged FP64 tensor-core (DMMA.8x8x4) GEMM -- the production path for aligned operands.
//
//   C[z][m][n] = sum_p sum_k opA(A[z,p][m][k]) * opB(B[z,p][k][n])          (contract: gemm.cu)
//
// * One cp.async.bulk.tensor (TMA) per operand per pipeline stage, issued by one elected thread, completion on
//   an mbarrier (expect-tx); a second mbarrier per stage (one arrival per warp) hands the slot back.  The
//   refill runs two iterations behind the consumers, so the issuing thread practically never waits.
// * Shared-memory tiles use the hardware 128-byte swizzle.  Fragment rows/columns are assigned to lanes through
//   small permutations (sigma / pos0 / rho below) chosen so that every LDS.64 (float64) or LDS.128 (complex128)
//   fragment load is bank-conflict free for both operand orientations; the permutation only renames which
//   output row/column a lane owns, which the epilogue undoes.
// * Either dimension of A and of B may be the contiguous one:
//     "KC"  k contiguous      : tensor map [k, r, p, z],           box [128 B of k, tile rows]
//     "RC"  row idx contiguous: tensor map [r_lo, k, r_hi, p, z],  box [128 B of r, BK, tile rows / r_lo]
//   complex128 is described to the TMA as float64 with a doubled innermost extent.
// * Out-of-range k (and rows of KC operands) are zero-filled by the TMA, so ragged extents need no copies.
#include "common.cuh"
#include <mutex>
#include <map>
#include <utility>
#include <cstdlib>

namespace qsb {

template <bool CPLX> struct TCfg;
template <> struct TCfg<false> {
    static constexpr int ES = 8, BM = 128, BN = 128, BK = 16;
    static constexpr int WARPS_M = 2, WARPS_N = 4, WM = 64, WN = 32;
    static constexpr int STAGES = 6;
    static constexpr int RLO = 16;                  // elements per 128-byte line
};
template <> struct TCfg<true> {
    static constexpr int ES = 16, BM = 128, BN = 64, BK = 8;
    static constexpr int WARPS_M = 4, WARPS_N = 2, WM = 32, WN = 32;
    static constexpr int STAGES = 8;
    static constexpr int RLO = 8;
};
constexpr int kTmaThreads = 256;

struct TArgs {
    int M, N, K, P;
    int tiles_m, tiles_n;
    int a_has_p, a_has_z, b_has_p, b_has_z;     // 0 when that stride is 0 (coordinate pinned to 0)
    char* C; long long c_ms, c_ns, c_zs;
    double sgnA, sgnB;
    // k-chunk gating (fused all-gather -> GEMM): the k range is ksplit equal chunks, chunk c becomes readable when
    // kflags[c] >= kepoch (set by the peer that pushed it); chunks are visited starting from chunk krot
    int raster;                 // tile rows per rasterisation group
    int ksplit, krot, kepoch, klocal;    // klocal: chunk krot is already in place, never wait for it
    int nsplit;                          // > 1: whole tiles are gated on the column block of B they read (kflags);
                                         // with ksplit > 1 too: flag kflags[kc * nsplit + column block] per k chunk
    int Zb;                              // batch count (needed for the n-outermost tile order of nsplit mode)
    const double* row_scale; int row_div;     // SCALED variant: C = row_scale[m / row_div] * col_scale[n % col_mod] * (A.B)
    const double* col_scale; int col_mod;
    const int* kflags;
};

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        // a lost completion must fail loudly, not hang -- but the producer of an arrival-gated matvec may itself be
        // waiting for a peer GPU, so the limit is generous (~60 s)
        if (clock64() - t0 > 120000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// lane -> fragment-row permutations (see header comment)
__device__ __forceinline__ int perm_sigma(int g) { return 2 * (g & 3) + (g >> 2); }               // f64, KC
__device__ __forceinline__ int perm_pos0(int g) { return (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2); }  // f64, RC
__device__ __forceinline__ int perm_rho(int g) { return (g >> 1) | ((g & 1) << 2); }               // c128

// index (inside the warp tile) of fragment i, lane group g
template <bool CPLX, bool KC>
__device__ __forceinline__ int frag_index(int i, int g) {
    if constexpr (CPLX) return 8 * i + perm_rho(g);
    else if constexpr (KC) return 8 * i + perm_sigma(g);
    else return 16 * (i >> 1) + 4 * (i & 1) + perm_pos0(g);
}

// Per-thread smem addressing of one operand: offset(i, s) = base + imm(i, s) + tab[sel(i, s)]
template <bool CPLX, bool KC> struct FragAddr {
    uint32_t base;
    uint32_t tab[4];
    __device__ __forceinline__ void init(int w0, int g, int t) {
        if constexpr (!CPLX && KC) {
            const int sg = perm_sigma(g);
            base = (uint32_t)((w0 + sg) * 128 + (t & 1) * 8);
#pragma unroll
            for (int s = 0; s < 4; ++s) tab[s] = (uint32_t)((((2 * s) ^ (sg & 6)) | ((t >> 1) ^ (sg & 1))) << 4);
        } else if constexpr (!CPLX && !KC) {
            const int ph = 4 * ((g >> 1) & 1) + (g >> 2);
            base = (uint32_t)((w0 / 16) * 2048 + t * 128 + (g & 1) * 8);
#pragma unroll
            for (int ih = 0; ih < 2; ++ih)
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) tab[ih * 2 + sl] = (uint32_t)(((ph | (2 * ih)) ^ (4 * sl + t)) << 4);
        } else if constexpr (CPLX && KC) {
            const int rg = perm_rho(g);
            base = (uint32_t)((w0 + rg) * 128);
            tab[0] = (uint32_t)((t ^ rg) << 4);
            tab[1] = (uint32_t)(((4 + t) ^ rg) << 4);
            tab[2] = tab[3] = 0;
        } else {
            const int rg = perm_rho(g);
            base = (uint32_t)((w0 / 8) * 1024 + t * 128);
            tab[0] = (uint32_t)((t ^ rg) << 4);
            tab[1] = (uint32_t)(((4 + t) ^ rg) << 4);
            tab[2] = tab[3] = 0;
        }
    }
    // byte offset of fragment i at k4-step s inside the operand tile
    __device__ __forceinline__ uint32_t at(int i, int s) const {
        if constexpr (!CPLX && KC) return base + (uint32_t)(i * 1024) + tab[s];
        else if constexpr (!CPLX && !KC) return base + (uint32_t)((i >> 1) * 2048 + s * 512) + tab[(i & 1) * 2 + (s & 1)];
        else if constexpr (CPLX && KC) return base + (uint32_t)(i * 1024) + tab[s];
        else return base + (uint32_t)(i * 1024 + s * 512) + tab[s];
    }
};

// ---- persistent tile schedule: data-parallel waves + a stream-K tail ------------------------------------
// The tile count is rarely a multiple of the SM count (chi = 2048, 8 GPUs: 128 tiles on 148 SMs), so the last
// partial wave would idle SMs.  The grid is one persistent CTA per SM.  Tiles [0, T_sk) -- the remainder wave plus
// one full wave -- are split in k ("stream-K"): their T_sk * ipt k-iterations are divided evenly over the CTAs;
// the remaining tiles are dealt round-robin, one whole tile per CTA per wave.  A CTA walks
// its stream-K range from the END backwards, so the piece it shares with the next-lower CTA comes last: the CTA that
// owns the final k-iteration of a tile ("finisher") adds the partial accumulators its lower-indexed neighbours wrote
// at the START of their own walk.  A CTA therefore only ever waits for lower-indexed CTAs (no co-residency assumption),
// and the partial sums are added in CTA order (deterministic).
struct Sched {
    long long u0, seg_end;      // stream-K unit range [u0, seg_end) still to do (units = k-iterations)
    int ipt;                    // k-iterations per tile
    int dp_next, dp_waves, G, c, T_sk;
    bool dp_first;
    __device__ __forceinline__ void init(long long U_sk, int ipt_, int T_sk_, int dp_waves_, int G_, int c_,
                                         bool dp_first_) {
        ipt = ipt_; T_sk = T_sk_; dp_waves = dp_waves_; G = G_; c = c_; dp_next = 0; dp_first = dp_first_;
        u0 = U_sk * c / G; seg_end = U_sk * (c + 1) / G;
    }
    // next segment: k-iterations [it0, it1) of tile `tile`
    // whole tiles (data-parallel waves)
    __device__ __forceinline__ bool next_dp(int& tile, int& it0, int& it1) {
        if (dp_next >= dp_waves) return false;
        tile = T_sk + dp_next * G + c; it0 = 0; it1 = ipt; ++dp_next;
        return true;
    }
    // this CTA's stream-K range, from its end backwards
    __device__ __forceinline__ bool next_sk(int& tile, int& it0, int& it1) {
        if (seg_end <= u0) return false;
        tile = (int)((seg_end - 1) / ipt);
        const long long tb = (long long)tile * ipt;
        const long long b = u0 > tb ? u0 : tb;
        it0 = (int)(b - tb); it1 = (int)(seg_end - tb);
        seg_end = b;
        return true;
    }
    // Stream-K first hides the fix-up wait behind the whole tiles that follow; the arrival-gated matvec instead
    // wants whole tiles first (they start at k = 0, i.e. with the block that is already local).
    __device__ __forceinline__ bool next(int& tile, int& it0, int& it1) {
        if (dp_first) return next_dp(tile, it0, it1) || next_sk(tile, it0, it1);
        return next_sk(tile, it0, it1) || next_dp(tile, it0, it1);
    }
};

// tile index -> (tm, tn, z).  Default: batch z outermost, grouped rasterisation inside (L2-friendly).  Column-gated
// mode: n outermost (then z, then m), so tiles are visited in the order their column block of B arrives.
template <bool GATED>
__device__ __forceinline__ void tile_map(int nsplit, int tiles_m, int tiles_n, int Zb, int raster, int tile,
                                         int& tm, int& tn, int& z) {
    if (GATED && nsplit > 1) {
        const int per_n = tiles_m * Zb;
        tn = tile / per_n;
        const int r = tile - tn * per_n;
        z = r / tiles_m;
        tm = r - z * tiles_m;
    } else {
        tile_coords(tile, tiles_m, tiles_n, tm, tn, z, raster);
    }
}

struct SkArgs {
    long long U_sk;             // stream-K units
    int T_sk, dp_waves, ipt;
    double* partial;            // [G][kAccPerThread][kTmaThreads]
    int* flags;                 // [G], flag == epoch when that CTA's partial is ready
    int epoch;
};
constexpr int kAccPerThread = 64;

// GATED instantiations carry the arrival-gating code (fused multi-GPU exchange, host upload); the plain ones do
// not: the kernel sits at the register limit and any extra live value perturbs the main loop (measured: 3 %).
// ACCUM instantiations add the tile to what C already holds (C += A.B): the identity-channel shortcut of the H_eff
// matvec pre-loads C with the panel whose environment slice is the identity and the GEMM runs over the other panels.
template <bool CPLX, bool A_KC, bool B_KC, int VARIANT>
__global__ void __launch_bounds__(kTmaThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TArgs g,
                const SkArgs sk) {
    using C = TCfg<CPLX>;
    constexpr bool GATED = VARIANT == 1, ACCUM = VARIANT == 2, SCALED = VARIANT == 3;
    constexpr int A_BYTES = C::BM * C::BK * C::ES;
    constexpr int B_BYTES = C::BN * C::BK * C::ES;
    constexpr int STAGE = A_BYTES + B_BYTES;
    constexpr int S = C::STAGES;
    constexpr int MI = C::WM / 8, NJ = C::WN / 8, KS = C::BK / 4;
    constexpr int ACC = CPLX ? 4 : 2;
    static_assert(MI * NJ * ACC == kAccPerThread, "partial-tile layout");

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + S * STAGE;            // S barriers, then S "empty" barriers
    const uint32_t bar_empty = bar_full + S * 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp / C::WARPS_N) * C::WM;
    const int wn0 = (warp % C::WARPS_N) * C::WN;
    const int cta = blockIdx.x, G = gridDim.x;
    const int ktiles = (g.K + C::BK - 1) / C::BK;

    if (tid == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, kTmaThreads / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    // ---- producer (thread 0): walks the same segment sequence as the consumers, S-2 stages ahead
    constexpr int KMUL = CPLX ? 2 : 1;       // innermost TMA coordinate is in float64 units
    Sched ps; ps.init(sk.U_sk, sk.ipt, sk.T_sk, sk.dp_waves, G, cta, GATED && g.ksplit > 1);
    int p_tile = -1, p_it = 0, p_it1 = 0, p_m0 = 0, p_n0 = 0, p_z = 0;
    int pj = 0;                              // CTA-local index of the next k-iteration to issue
    int p_chunk_ok = -1;                     // last gated k chunk known to have arrived
    bool p_more = true;
    auto produce = [&]() {                   // issue one k-iteration (if any is left); thread 0 only
        if (p_it == p_it1) {
            if (!p_more || !ps.next(p_tile, p_it, p_it1)) { p_more = false; return; }
            int tm, tn;
            tile_map<GATED>(g.nsplit, g.tiles_m, g.tiles_n, g.Zb, g.raster, p_tile, tm, tn, p_z);
            p_m0 = tm * C::BM; p_n0 = tn * C::BN;
            if (GATED && g.nsplit > 1 && g.ksplit <= 1) {     // column-block gating: this tile reads one block of B, wait for it once
                const int c = (int)((long long)p_n0 * g.nsplit / g.N);
                if (c != p_chunk_ok) {
                    int v;
                    const long long t0 = clock64();
                    do {
                        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(g.kflags + c) : "memory");
                        if (v < g.kepoch && clock64() - t0 > 100000000000LL) __trap();
                    } while (v < g.kepoch);
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    p_chunk_ok = c;
                }
            }
        }
        const int slot = pj % S;
        if (pj >= S) mbar_wait_spin(bar_empty + 8 * slot, ((pj / S) + 1) & 1);
        int p = p_it / ktiles, k0 = (p_it - p * ktiles) * C::BK;
        if (GATED && g.ksplit > 1) {         // P == 1 here: visit the k chunks in rotated order, gated on arrival
            const int per = ktiles / g.ksplit;
            const int c = p_it / per;
            int kc = c + g.krot; if (kc >= g.ksplit) kc -= g.ksplit;
            // grid gating (nsplit > 1 as well): theta arrives in ksplit row blocks x nsplit column blocks, one flag each
            const int fi = g.nsplit > 1 ? kc * g.nsplit + (int)((long long)p_n0 * g.nsplit / g.N) : kc;
            if (fi != p_chunk_ok && !(g.klocal && kc == g.krot)) {     // the local block is never gated
                int v;
                const long long t0 = clock64();
                do {
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(g.kflags + fi) : "memory");
                    if (v < g.kepoch && clock64() - t0 > 100000000000LL) __trap();     // ~50 s: peer never pushed
                } while (v < g.kepoch);
                asm volatile("fence.proxy.async.global;" ::: "memory");   // peer-written data -> TMA (async proxy) reads
                p_chunk_ok = fi;
            }
            p = 0;
            k0 = (kc * per + (p_it - c * per)) * C::BK;
        }
        const uint32_t sa = sbase + slot * STAGE, sb = sa + A_BYTES, bar = bar_full + 8 * slot;
        mbar_expect_tx(bar, STAGE);
        if constexpr (A_KC) tma_load_4d(sa, &mapA, bar, k0 * KMUL, p_m0, g.a_has_p ? p : 0, g.a_has_z ? p_z : 0);
        else tma_load_5d(sa, &mapA, bar, 0, k0, p_m0 / C::RLO, g.a_has_p ? p : 0, g.a_has_z ? p_z : 0);
        if constexpr (B_KC) tma_load_4d(sb, &mapB, bar, k0 * KMUL, p_n0, g.b_has_p ? p : 0, g.b_has_z ? p_z : 0);
        else tma_load_5d(sb, &mapB, bar, 0, k0, p_n0 / C::RLO, g.b_has_p ? p : 0, g.b_has_z ? p_z : 0);
        ++p_it; ++pj;
    };
    if (tid == 0) {
        for (int j = 0; j < S - 2; ++j) produce();
    }

    FragAddr<CPLX, A_KC> fa; fa.init(wm0, g8, t4);
    FragAddr<CPLX, B_KC> fb; fb.init(wn0, g8, t4);

    Sched cs; cs.init(sk.U_sk, sk.ipt, sk.T_sk, sk.dp_waves, G, cta, GATED && g.ksplit > 1);
    int tile, it0, it1;
    int cj = 0;                              // CTA-local index of the k-iteration being consumed
    double* my_partial = sk.partial + (size_t)cta * kAccPerThread * kTmaThreads + tid;

    while (cs.next(tile, it0, it1)) {
        double acc[MI][NJ][ACC];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int q = 0; q < ACC; ++q) acc[i][j][q] = 0.0;

        for (int it = it0; it < it1; ++it, ++cj) {
            if (tid == 0) produce();
            const int slot = cj % S;
            mbar_wait_spin(bar_full + 8 * slot, (cj / S) & 1);
            const unsigned char* sa = smem + slot * STAGE;
            const unsigned char* sb = sa + A_BYTES;
            if constexpr (!CPLX) {
                double a[2][MI], b[2][NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[0][i] = *reinterpret_cast<const double*>(sa + fa.at(i, 0));
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[0][j] = *reinterpret_cast<const double*>(sb + fb.at(j, 0));
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const int cur = kk & 1, nxt = cur ^ 1;
                    if (kk + 1 < KS) {
#pragma unroll
                        for (int i = 0; i < MI; ++i) a[nxt][i] = *reinterpret_cast<const double*>(sa + fa.at(i, kk + 1));
#pragma unroll
                        for (int j = 0; j < NJ; ++j) b[nxt][j] = *reinterpret_cast<const double*>(sb + fb.at(j, kk + 1));
                    }
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
                }
            } else {
                double2 a[2][MI], b[2][NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[0][i] = *reinterpret_cast<const double2*>(sa + fa.at(i, 0));
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[0][j] = *reinterpret_cast<const double2*>(sb + fb.at(j, 0));
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const int cur = kk & 1, nxt = cur ^ 1;
                    if (kk + 1 < KS) {
#pragma unroll
                        for (int i = 0; i < MI; ++i) a[nxt][i] = *reinterpret_cast<const double2*>(sa + fa.at(i, kk + 1));
#pragma unroll
                        for (int j = 0; j < NJ; ++j) b[nxt][j] = *reinterpret_cast<const double2*>(sb + fb.at(j, kk + 1));
                    }
                    double ai[MI], nai[MI], bi[NJ];
#pragma unroll
                    for (int i = 0; i < MI; ++i) { ai[i] = a[cur][i].y * g.sgnA; nai[i] = -ai[i]; }
#pragma unroll
                    for (int j = 0; j < NJ; ++j) bi[j] = b[cur][j].y * g.sgnB;
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
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
        }

        // ---- segment epilogue
        if (it1 < sk.ipt) {
            // contributor: this CTA does not own the tile's last k-iteration -> publish the partial accumulators
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int q = 0; q < ACC; ++q)
                        __stcg(my_partial + (size_t)((i * NJ + j) * ACC + q) * kTmaThreads, acc[i][j][q]);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(sk.flags + cta), "r"(sk.epoch) : "memory");
            }
            continue;
        }
        if (it0 > 0) {
            // finisher: add the partials of the lower-indexed CTAs that cover k-iterations [0, it0) of this tile
            const long long tile_begin = (long long)tile * sk.ipt;
            for (int cc = cta - 1; cc >= 0 && sk.U_sk * (cc + 1) / G > tile_begin; --cc) {
                if (tid == 0) {
                    int v;
                    const long long t0 = clock64();
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(sk.flags + cc) : "memory");
                        if (v != sk.epoch && clock64() - t0 > 100000000000LL) __trap();
                    } while (v != sk.epoch);
                }
                __syncthreads();
                const double* src = sk.partial + (size_t)cc * kAccPerThread * kTmaThreads + tid;
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int q = 0; q < ACC; ++q)
                            acc[i][j][q] += __ldcg(src + (size_t)((i * NJ + j) * ACC + q) * kTmaThreads);
            }
        }
        // store: lane owns rows frag_index<A>(i, g8) and columns frag_index<B>(j, 2*t4), frag_index<B>(j, 2*t4+1)
        int tm, tn, z;
        tile_map<GATED>(g.nsplit, g.tiles_m, g.tiles_n, g.Zb, g.raster, tile, tm, tn, z);
        const int m0 = tm * C::BM, n0 = tn * C::BN;
        char* Cz = g.C + g.c_zs * z * C::ES;
        bool vec_ok = false;
        if constexpr (!CPLX && !B_KC)        // RC columns come in adjacent pairs -> 16-byte stores
            vec_ok = (g.c_ns == 1) && ((g.c_ms & 1) == 0) && ((g.c_zs & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + wm0 + frag_index<CPLX, A_KC>(i, g8);
            if (m >= g.M) continue;
            double rsv = 1.0;
            if constexpr (SCALED) { if (g.row_scale) rsv = g.row_scale[m / g.row_div]; }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int c0 = n0 + wn0 + frag_index<CPLX, B_KC>(j, 2 * t4);
                const int c1 = n0 + wn0 + frag_index<CPLX, B_KC>(j, 2 * t4 + 1);
                char* row = Cz + g.c_ms * m * C::ES;
                if constexpr (SCALED) {          // fused Schmidt scaling of the theta build
                    const double s0 = rsv * ((g.col_scale && c0 < g.N) ? g.col_scale[c0 % g.col_mod] : 1.0);
                    const double s1 = rsv * ((g.col_scale && c1 < g.N) ? g.col_scale[c1 % g.col_mod] : 1.0);
                    acc[i][j][0] *= s0; acc[i][j][1] *= s1;
                    if constexpr (CPLX) { acc[i][j][2] *= s0; acc[i][j][3] *= s1; }
                }
                if constexpr (!CPLX) {
                    if (vec_ok && c1 < g.N) {        // c1 == c0 + 1 for the RC permutation
                        double2* dst = reinterpret_cast<double2*>(row + (long long)c0 * C::ES);
                        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                        if constexpr (ACCUM) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                        *dst = v;
                    } else {
                        if (c0 < g.N) {
                            double* dst = reinterpret_cast<double*>(row + g.c_ns * c0 * C::ES);
                            *dst = ACCUM ? acc[i][j][0] + *dst : acc[i][j][0];
                        }
                        if (c1 < g.N) {
                            double* dst = reinterpret_cast<double*>(row + g.c_ns * c1 * C::ES);
                            *dst = ACCUM ? acc[i][j][1] + *dst : acc[i][j][1];
                        }
                    }
                } else {
                    if (c0 < g.N) {
                        double2* dst = reinterpret_cast<double2*>(row + g.c_ns * c0 * C::ES);
                        double2 v = make_double2(acc[i][j][0], acc[i][j][2]);
                        if constexpr (ACCUM) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                        *dst = v;
                    }
                    if (c1 < g.N) {
                        double2* dst = reinterpret_cast<double2*>(row + g.c_ns * c1 * C::ES);
                        double2 v = make_double2(acc[i][j][1], acc[i][j][3]);
                        if constexpr (ACCUM) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                        *dst = v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor-map construction and dispatch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    });
    return fn;
}

struct Operand {
    const void* ptr; long long rs, ks, ps, zs;   // element strides (rs: row index m or n)
    int rows, has_p, has_z;
};

// Can this operand be described to the TMA?  (16-byte aligned base and strides; RC rows a multiple of a line)
static bool tma_eligible(const Operand& o, bool cplx, bool kc, int K, int P, int Z) {
    const long long es = cplx ? 16 : 8;
    if (!aligned16(o.ptr)) return false;
    auto ok = [&](long long st) { return st > 0 && (st * es) % 16 == 0 && st * es < (1LL << 40); };
    if (kc) { if (o.rows > 1 && !ok(o.rs)) return false; }
    else {
        const int rlo = cplx ? 8 : 16;
        if (o.rows % rlo != 0) return false;
        if (K > 1 && !ok(o.ks)) return false;
    }
    if (o.has_p && P > 1 && !ok(o.ps)) return false;
    if (o.has_z && Z > 1 && !ok(o.zs)) return false;
    return true;
}

static bool make_map(CUtensorMap* map, const Operand& o, bool cplx, bool kc, int tile_rows, int BK, int K, int P,
                     int Z) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const long long es = cplx ? 16 : 8;
    const int kmul = cplx ? 2 : 1;
    const int rlo = cplx ? 8 : 16;
    cuuint64_t dims[5]; cuuint64_t strides[4]; cuuint32_t box[5]; cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const cuuint64_t np = (o.has_p && P > 1) ? (cuuint64_t)P : 1, nz = (o.has_z && Z > 1) ? (cuuint64_t)Z : 1;
    // strides of extent-1 dimensions are never used for addressing, but must still be valid (16-byte multiples)
    auto st = [&](long long s, bool used) { return (cuuint64_t)((used ? s : 2) * es); };
    int rank;
    if (kc) {
        rank = 4;
        dims[0] = (cuuint64_t)K * kmul; dims[1] = (cuuint64_t)o.rows; dims[2] = np; dims[3] = nz;
        strides[0] = st(o.rs, o.rows > 1); strides[1] = st(o.ps, np > 1); strides[2] = st(o.zs, nz > 1);
        box[0] = (cuuint32_t)(BK * kmul); box[1] = (cuuint32_t)tile_rows; box[2] = 1; box[3] = 1;
    } else {
        rank = 5;
        dims[0] = (cuuint64_t)rlo * kmul; dims[1] = (cuuint64_t)K; dims[2] = (cuuint64_t)(o.rows / rlo);
        dims[3] = np; dims[4] = nz;
        strides[0] = st(o.ks, K > 1); strides[1] = (cuuint64_t)(rlo * es); strides[2] = st(o.ps, np > 1);
        strides[3] = st(o.zs, nz > 1);
        box[0] = (cuuint32_t)(rlo * kmul); box[1] = (cuuint32_t)BK; box[2] = (cuuint32_t)(tile_rows / rlo);
        box[3] = 1; box[4] = 1;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(o.ptr), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// stream-K scratch (partial tiles + flags + epoch), owned by the library: one per (device, stream), so GEMMs issued
// on different streams of one device never share partial-tile storage (re-entrant per (device, stream)).
// Launches on ONE stream are ordered, which is all a scratch needs.  Entries live for the life of the process.
struct SkScratch { double* partial = nullptr; int* flags = nullptr; int sms = 0; int epoch = 0; };
static std::mutex g_sk_mutex;
static std::map<std::pair<int, cudaStream_t>, SkScratch> g_sk;

static int sk_scratch(SkScratch** out, cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return QSB_ERR_DEVICE;
    std::lock_guard<std::mutex> lock(g_sk_mutex);
    SkScratch& s = g_sk[std::make_pair(dev, stream)];
    if (!s.partial) {
        int rc = check_cuda(cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev));
        if (rc) return rc;
        rc = check_cuda(cudaMalloc(&s.partial, (size_t)s.sms * kAccPerThread * kTmaThreads * sizeof(double)));
        if (rc) return rc;
        rc = check_cuda(cudaMalloc(&s.flags, (size_t)s.sms * sizeof(int)));
        if (rc) return rc;
        rc = check_cuda(cudaMemset(s.flags, 0, (size_t)s.sms * sizeof(int)));
        if (rc) return rc;
    }
    *out = &s;
    return QSB_OK;
}

template <bool CPLX, bool A_KC, bool B_KC, int VARIANT>
static int launch_tma_g(const CUtensorMap& ma, const CUtensorMap& mb, const TArgs& t, int Z, cudaStream_t stream) {
    using C = TCfg<CPLX>;
    auto kern = gemm_tma_kernel<CPLX, A_KC, B_KC, VARIANT>;
    constexpr int SMEM = C::STAGES * (C::BM + C::BN) * C::BK * C::ES + 2 * C::STAGES * 8 + 1024;
    static bool configured[64] = {};         // per instantiation and device (the attribute is per device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return QSB_ERR_DEVICE;
    if (!configured[dev]) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        if (rc) return rc;
        configured[dev] = true;
    }
    SkScratch* sc = nullptr;
    int rc = sk_scratch(&sc, stream);
    if (rc) return rc;
    const long long tiles = (long long)t.tiles_m * t.tiles_n * Z;
    if (tiles > 2147483647LL) return QSB_ERR_SHAPE;
    const int ipt = ((t.K + C::BK - 1) / C::BK) * t.P;
    // schedule: one persistent CTA per SM; whole waves data-parallel, the remainder (+ one wave) split in k
    constexpr int kMinIters = 16;             // do not split below this many k-iterations per CTA
    int G = sc->sms;
    SkArgs sk;
    sk.ipt = ipt;
    if (tiles < G) {
        long long by_work = tiles * ipt / kMinIters;
        if (by_work < tiles) by_work = tiles;
        if (by_work < G) G = (int)by_work;
    }
    const long long full = tiles / G, rem = tiles % G;
    if (rem == 0) { sk.T_sk = 0; sk.dp_waves = (int)full; }
    else { sk.dp_waves = (int)(full > 0 ? full - 1 : 0); sk.T_sk = (int)(tiles - (long long)sk.dp_waves * G); }
    sk.U_sk = (long long)sk.T_sk * ipt;
    sk.partial = sc->partial; sk.flags = sc->flags;
    sk.epoch = ++sc->epoch;
    if (sc->epoch == 0x7fffffff) sc->epoch = 0;
    kern<<<(unsigned)G, kTmaThreads, SMEM, stream>>>(ma, mb, t, sk);
    ++g_count_tma;
    return check_launch();
}

template <bool CPLX, bool A_KC, bool B_KC>
static int launch_tma(const CUtensorMap& ma, const CUtensorMap& mb, const TArgs& t, int Z, cudaStream_t stream,
                      bool accumulate) {
    if (t.ksplit > 1 || t.nsplit > 1) {
        if (accumulate || t.row_scale || t.col_scale) return QSB_ERR_ARG;   // no caller needs a gated accumulating / scaled GEMM
        return launch_tma_g<CPLX, A_KC, B_KC, 1>(ma, mb, t, Z, stream);
    }
    if (t.row_scale || t.col_scale) {
        if (accumulate) return QSB_ERR_ARG;
        return launch_tma_g<CPLX, A_KC, B_KC, 3>(ma, mb, t, Z, stream);
    }
    if (accumulate) return launch_tma_g<CPLX, A_KC, B_KC, 2>(ma, mb, t, Z, stream);
    return launch_tma_g<CPLX, A_KC, B_KC, 0>(ma, mb, t, Z, stream);
}

int gemm_tma_try(const GemmArgs& g, cudaStream_t stream, bool* used) {
    *used = false;
    const bool cplx = g.cplx != 0;
    if (g.K < 1 || g.P < 1) return QSB_OK;
    Operand a{g.A, g.a_rs, g.a_ks, g.a_ps, g.a_zs, g.M, g.a_ps != 0, g.a_zs != 0};
    Operand b{g.B, g.b_rs, g.b_ks, g.b_ps, g.b_zs, g.N, g.b_ps != 0, g.b_zs != 0};
    // orientation: prefer k-contiguous when both strides are 1 (extent-1 dimensions)
    const bool a_kc = (a.ks == 1 || g.K == 1), b_kc = (b.ks == 1 || g.K == 1);
    if (!a_kc && a.rs != 1) return QSB_OK;
    if (!b_kc && b.rs != 1) return QSB_OK;
    if (a_kc && g.K == 1) a.ks = 1;
    if (b_kc && g.K == 1) b.ks = 1;
    if (!tma_eligible(a, cplx, a_kc, g.K, g.P, g.Z) || !tma_eligible(b, cplx, b_kc, g.K, g.P, g.Z)) return QSB_OK;
    if (!aligned16(g.C)) return QSB_OK;
    const int BM = cplx ? TCfg<true>::BM : TCfg<false>::BM, BN = cplx ? TCfg<true>::BN : TCfg<false>::BN;
    const int BK = cplx ? TCfg<true>::BK : TCfg<false>::BK;
    CUtensorMap ma, mb;
    if (!make_map(&ma, a, cplx, a_kc, BM, BK, g.K, g.P, g.Z)) return QSB_OK;
    if (!make_map(&mb, b, cplx, b_kc, BN, BK, g.K, g.P, g.Z)) return QSB_OK;

    TArgs t;
    t.M = g.M; t.N = g.N; t.K = g.K; t.P = g.P;
    t.tiles_m = (g.M + BM - 1) / BM; t.tiles_n = (g.N + BN - 1) / BN;
    t.a_has_p = (a.has_p && g.P > 1); t.a_has_z = (a.has_z && g.Z > 1);
    t.b_has_p = (b.has_p && g.P > 1); t.b_has_z = (b.has_z && g.Z > 1);
    t.C = (char*)g.C; t.c_ms = g.c_ms; t.c_ns = g.c_ns; t.c_zs = g.Z > 1 ? g.c_zs : 0;
    t.sgnA = g.conjA ? -1.0 : 1.0; t.sgnB = g.conjB ? -1.0 : 1.0;
    {
        static int raster_env = -1;          // experiment knob: QSB_RASTER_GROUP overrides the default group size
        if (raster_env < 0) { const char* e = getenv("QSB_RASTER_GROUP"); raster_env = e ? atoi(e) : 0; }
        t.raster = raster_env > 0 ? raster_env : kRasterGroup;
    }
    t.ksplit = 1; t.krot = 0; t.kepoch = 0; t.klocal = 0; t.kflags = nullptr; t.nsplit = 1; t.Zb = g.Z;
    t.row_scale = g.row_scale; t.row_div = g.row_div > 0 ? g.row_div : 1;
    t.col_scale = g.col_scale; t.col_mod = g.col_mod > 0 ? g.col_mod : 1;
    if (g.ksplit > 1) {
        const int ktiles = (g.K + BK - 1) / BK;
        if (g.P != 1 || g.K % BK != 0 || ktiles % g.ksplit != 0 || !g.kflags) return QSB_ERR_STRIDE;
        t.ksplit = g.ksplit; t.krot = g.krot % g.ksplit; t.kepoch = g.kepoch; t.kflags = g.kflags; t.klocal = g.klocal;
    }
    if (g.nsplit > 1) {                  // tiles must not straddle column blocks
        if (!g.kflags || g.N % g.nsplit != 0 || (g.N / g.nsplit) % BN != 0) return QSB_ERR_STRIDE;
        t.nsplit = g.nsplit; t.kepoch = g.kepoch; t.kflags = g.kflags;
    }
    int rc;
    const bool acc = g.accumulate != 0;
    if (cplx) {
        if (a_kc) rc = b_kc ? launch_tma<true, true, true>(ma, mb, t, g.Z, stream, acc) : launch_tma<true, true, false>(ma, mb, t, g.Z, stream, acc);
        else rc = b_kc ? launch_tma<true, false, true>(ma, mb, t, g.Z, stream, acc) : launch_tma<true, false, false>(ma, mb, t, g.Z, stream, acc);
    } else {
        if (a_kc) rc = b_kc ? launch_tma<false, true, true>(ma, mb, t, g.Z, stream, acc) : launch_tma<false, true, false>(ma, mb, t, g.Z, stream, acc);
        else rc = b_kc ? launch_tma<false, false, true>(ma, mb, t, g.Z, stream, acc) : launch_tma<false, false, false>(ma, mb, t, g.Z, stream, acc);
    }
    if (rc) return rc;
    *used = true;
    return QSB_OK;
}

}  // namespace qsb

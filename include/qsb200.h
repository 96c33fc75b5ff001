/*
 * qsb200.h -- C ABI of the B200-native two-site DMRG inner loop.
 *
 * Drop-in boundary for the hot path of ToelUl/quantum-simulation (pure Python;
 * file:line below are relative to /root/reference/quantum_simulation/):
 *
 *   qsb_heff_apply        replaces ApplyMPO.forward            dmrg.py:178-248
 *                         (the ncon_torch contraction path     ncon_torch.py:111-154)
 *   qsb_env_left          replaces LeftEnvUpdate.forward       dmrg.py:254-278
 *   qsb_env_right         replaces RightEnvUpdate.forward      dmrg.py:284-308
 *   qsb_lz_*              replace the vector ops of
 *                         _LanczosEngine.forward               eigensolver/lanczos_solver.py:228-295
 *                         (vdot :241, axpy :245-247, PRO :250-256, norm :259,
 *                          copy/scale :230,268-270, Ritz combine :293-294,
 *                          _normalize_to :137-155)
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host; no torch types anywhere.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered,
 *     nothing synchronises the host.
 *   - every function returns a qsb_status (0 = ok) and never throws; the
 *     Python layer maps the codes to the reference's exception types.
 *   - dtype codes: QSB_F64 / QSB_C128 for the contractions; the Lanczos vector
 *     ops additionally take QSB_F32 / QSB_C64 (the reference's test matrix,
 *     eigensolver/test_lanczos.py:10).  complex = interleaved (re, im).
 *   - tensor index conventions are the reference's (dmrg.py:188-207):
 *       theta[a, s1, s2, b]   flat, row-major
 *       L[w, a_out, a_in]     element strides given (any of the two layouts
 *                             the reference produces, or arbitrary -> copied)
 *       M[w_left, w_right, s_out, s_in]   contiguous
 *       R[w, b_out, b_in]     element strides given
 *       MPS site tensor [chi_l, d, chi_r], element strides given
 */
#ifndef QSB200_H
#define QSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QSB_OK = 0,
    QSB_ERR_SHAPE = 1,      /* inconsistent extents (reference: AssertionError dmrg.py:233-237) */
    QSB_ERR_DTYPE = 2,
    QSB_ERR_STRIDE = 3,
    QSB_ERR_DEVICE = 4,     /* no usable sm_100 device / wrong device */
    QSB_ERR_WORKSPACE = 5,  /* workspace pointer NULL or too small */
    QSB_ERR_LAUNCH = 6,     /* CUDA launch / runtime failure (see qsb_last_cuda_error) */
    QSB_ERR_ARG = 7
} qsb_status;

typedef enum { QSB_F64 = 0, QSB_C128 = 1, QSB_F32 = 2, QSB_C64 = 3 } qsb_dtype;

/* ---- library info ------------------------------------------------------ */
int         qsb_version(void);
const char* qsb_status_string(int status);
int         qsb_last_cuda_error(void);        /* cudaError_t of the last QSB_ERR_LAUNCH */
/* counters of GEMM launches by staging path since load / last reset:
   out[0] = TMA-staged launches, out[1] = cp.async-staged (ragged/unaligned) launches,
   out[2] = all other kernels launched by the library */
void        qsb_launch_counters(uint64_t out[3], int reset);

/* ---- H_eff . theta  (dmrg.py:178-248) ---------------------------------- */
typedef struct {
    int dtype;                                  /* QSB_F64 | QSB_C128 */
    int chi_l_out, chi_l_in, chi_r_out, chi_r_in;
    int d1_out, d1_in, d2_out, d2_in;
    int W_l, W_m, W_r;
    const void* theta;                          /* (chi_l_in, d1_in, d2_in, chi_r_in) contiguous */
    const void* L;  int64_t L_stride[3];        /* (W_l, chi_l_out, chi_l_in), element strides */
    const void* M1;                             /* (W_l, W_m, d1_out, d1_in) contiguous */
    const void* M2;                             /* (W_m, W_r, d2_out, d2_in) contiguous */
    const void* R;  int64_t R_stride[3];        /* (W_r, chi_r_out, chi_r_in), element strides */
    void*       out;                            /* (chi_l_out, d1_out, d2_out, chi_r_out) contiguous */
    void*       workspace; size_t workspace_bytes;
    /* Fused all-gather -> matvec (multi-GPU, optional; n_chunks <= 1 disables it).  theta is the gathered vector
       whose chi_l_in rows arrive in n_chunks equal blocks, block c pushed by rank c (qsb_push_shard).  Step 1 then
       walks its k range block by block starting with block chunk_rot (the caller's own, already local) and reads
       block c only once chunk_flags[c] >= chunk_epoch, so the transfer overlaps the first tiles. */
    int n_chunks, chunk_rot, chunk_epoch;
    const int* chunk_flags;
    int chunk_axis;         /* 0: the blocks are ROW blocks of theta (the k range of step 1; multi-GPU exchange).
                               1: the blocks are COLUMN blocks of theta viewed as (chi_l_in, d1*d2*chi_r): an output tile
                                  of step 1 then depends on ONE block only, so tiles start as soon as their block has
                                  landed (host upload with qsb_upload_columns); chunk_rot / chunk_local are ignored.
                               2: a GRID of n_chunks row blocks x chunk_cols column blocks, flag index
                                  row_block * chunk_cols + column_block: every rank uploads its row shard of a HOST-resident
                                  theta in column blocks and forwards each block to its peers (qsb_upload_columns +
                                  qsb_push_columns_dma), and tiles of column block j start once (r, j) is there for the
                                  row block r they are reading. */
    int chunk_local;        /* 1: block chunk_rot is already in place (stream-ordered before this call), do not gate it.
                               Required when that block is copied on the SAME device: such a copy may need SMs, which the
                               persistent GEMM would be holding while it waits for the flag. */
    /* Identity-channel shortcut (MPO-sparsity-aware block skipping in the chi^3 GEMMs).  For finite-state-machine
       MPOs (mpo/auto_mpo/finite_state_machine.py:662-700) the "nothing has happened yet" channel of the left
       environment and the "everything is done" channel of the right one are identity matrices whenever the MPS is in
       canonical form (boundary ones(W,1,1), dmrg.py:349-351, times isometries).  The caller VERIFIES this numerically
       (qsb_env_identity_deviation) and passes 1 + the channel index (0 = no such channel).  Step 1 then skips the GEMM
       with L[w] (T1[w] is theta itself), step 4 skips the panel of R[w] (its T3 slice goes straight to `out`).
       Requires chi_l_out <= chi_l_in resp. chi_r_out == chi_r_in; ignored otherwise. */
    int L_identity;         /* 1 + w with L[w][a'][a] == delta(a, a' + L_row0) */
    int R_identity;         /* 1 + w with R[w][b'][b] == delta(b, b') */
    int L_row0;             /* first row of the bra index held by a row-sharded L (multi-GPU), else 0 */
    int chunk_cols;         /* chunk_axis == 2: number of column blocks per row block (else ignored) */
} qsb_heff_desc;

int qsb_heff_workspace_bytes(const qsb_heff_desc* d, size_t* bytes);
int qsb_heff_apply(const qsb_heff_desc* d, void* stream);
/* dense ncon-order flop count of one matvec: the "effective" figure */
double qsb_heff_flops(const qsb_heff_desc* d);
/* flops the two GEMMs of THIS call execute (identity channels skipped) plus the dense mix count: the figure a
   tensor-pipe utilisation must be computed from (never credit skipped work) */
double qsb_heff_flops_executed(const qsb_heff_desc* d);
/* maxdev[w] = max over (r, c) of | env[w][r][c] - delta(c, r + row0) |  for every channel w < W of an environment
   (W, rows, cols) with element strides: the numerical test behind L_identity / R_identity.  maxdev: W doubles on the
   device, written stream-ordered. */
int qsb_env_identity_deviation(int dtype, int W, int rows, int cols, const void* env, const int64_t stride[3],
                               int row0, double* maxdev, void* stream);
/* Profiling hook: with count == 4 every following qsb_heff_apply records these cudaEvent_t on its stream:
   [0] before step 1, [1] after step 1 (GEMM), [2] after the MPO mix, [3] after step 4 (GEMM).
   count == 0 clears the hook.  Used by bench.py to time the dominant kernel inside the timed region. */
int qsb_heff_set_stage_events(void* const* events, int count);

/* Push one row shard of a vector into every rank's gathered buffer over NVLink and publish it:
   dst[r] (r < world) is the address of THIS rank's block inside rank r's buffer (peer-mapped memory, 16-byte
   aligned), flag[r] the address of THIS rank's arrival flag on rank r; after all stores, *flag[r] = epoch
   (system-scope release).  `ticket` is a zero-initialised device int owned by the caller. */
int qsb_push_shard(const void* src, size_t bytes, void* const* dst, int* const* flag, int world, int epoch,
                   void* ticket, void* stream);
/* Same contract on the copy engines: for each of the `count` destinations, in the given order, one peer
   cudaMemcpyAsync followed by one stream-ordered 32-bit flag write (cuStreamWriteValue32).  No SM is used, so on a
   side stream it overlaps the persistent GEMM of the matvec that consumes the blocks.  `src` may also be PINNED HOST
   memory (cudaMemcpyDefault): the same gating then overlaps a host->device upload of theta with the step-1 GEMM. */
int qsb_push_shard_dma(const void* src, size_t bytes, void* const* dst, int* const* flag, int count, int epoch,
                       void* stream);

/* Upload a row-major (rows x row_elems) matrix from PINNED HOST memory in n_chunks column blocks on the copy engines
   (cudaMemcpy2DAsync), each followed by a stream-ordered flag write flags[c] = epoch; pairs with chunk_axis = 1. */
int qsb_upload_columns(const void* src_host, void* dst, int64_t rows, int64_t row_elems, int elem_bytes, int n_chunks,
                       int* flags, int epoch, void* stream);

/* Forward a (rows x row_elems) row-major device block to `count` peers in n_chunks column blocks on the copy engines: for
   each column block c, first a stream-ordered wait until ready_flags[c] >= epoch (cuStreamWaitValue32: the block has been
   uploaded into `src` by qsb_upload_columns on another stream), then per destination one cudaMemcpy2DAsync into dst[i]
   (same pitch) and a flag write flag[i][c] = epoch.  Pairs with chunk_axis = 2. */
int qsb_push_columns_dma(const void* src, int64_t rows, int64_t row_elems, int elem_bytes, int n_chunks,
                         const int* ready_flags, void* const* dst, int* const* flag, int count, int epoch, void* stream);

/* ---- environment updates (dmrg.py:254-308) ----------------------------- */
typedef struct {
    int dtype;
    int W_in, W_out;        /* MPO bond on the old-env side / new-env side */
    int d;                  /* physical dimension */
    int chi_in, chi_out;    /* MPS bond on the old-env side / new-env side */
    const void* env;  int64_t env_stride[3];    /* left : (W_in, chi_bra or chi_in, chi_in); right: (W_in, chi_in, chi_in) */
    const void* M;                              /* left : (W_in, W_out, d, d); right: (W_out, W_in, d, d) */
    const void* site; int64_t site_stride[3];   /* left : A (chi_in, d, chi_out); right: B (chi_out, d, chi_in) */
    void*       out;                            /* left : (W_out, chi_out, chi_out); right: (W_out, chi_bra or chi_out, chi_out); contiguous */
    void*       workspace; size_t workspace_bytes;
    /* Row-block form used by the multi-GPU sharding over the bra index (chi_bra == 0: plain update).
       left : env holds only rows [lo, lo+chi_bra) of its bra index and site_bra = A[lo:lo+chi_bra] -> out is this
              rank's PARTIAL SUM of the full new environment (reduce-scatter across ranks follows);
       right: site_bra = B[lo:lo+chi_bra] -> out holds rows [lo, lo+chi_bra) of the new environment. */
    int chi_bra;
    const void* site_bra; int64_t site_bra_stride[3];
    /* 1 + w of an identity channel of `env` (verified by the caller, see qsb_heff_desc.L_identity): the first GEMM
       with that slice is replaced by a (conjugating) copy of the site tensor.  Plain form only (chi_bra == 0). */
    int env_identity;
} qsb_env_desc;

int qsb_env_workspace_bytes(const qsb_env_desc* d, size_t* bytes);
int qsb_env_left(const qsb_env_desc* d, void* stream);
int qsb_env_right(const qsb_env_desc* d, void* stream);
double qsb_env_flops(const qsb_env_desc* d);

/* ---- Lanczos vector ops (lanczos_solver.py:137-155, 228-295) ------------
 * Device scalars live in an array `scal` of (re, im) double pairs owned by
 * the caller; kernels read coefficients from it and write results to it, so
 * a whole Krylov step is stream-ordered without host round trips.
 * `scratch` is a caller-owned device buffer of qsb_lz_scratch_bytes() bytes
 * (block partial sums + a ticket counter, zero-initialised once).
 * Reductions are two-stage with a fixed summation order: bit-reproducible.
 */
size_t qsb_lz_scratch_bytes(int max_vectors);

/* flag bits shared by the reductions below */
#define QSB_LZ_REAL   1   /* keep only the real part (alpha, lanczos_solver.py:241) */
#define QSB_LZ_SQRT   2   /* qsb_lz_dot: store sqrt(re) (norm, :259, with X == w, m == 1) */
#define QSB_LZ_NOSQRT 4   /* axpy_norm / combine: store the SUM OF SQUARES instead of the norm -- the partial
                             result of one row shard, to be all-reduced across GPUs before the square root */

/* scal[out_idx + i] = <X_i, w>  (conjugate-linear in X), i < m; X_i = X + i*ldx. */
int qsb_lz_dot(int dtype, int64_t n, int m, const void* X, int64_t ldx, const void* w,
               double* scal, int out_idx, int flags, void* scratch, void* stream);

/* w <- w - sum_{i<m} scal[coef_idx + i] * X_i ;  scal[norm_idx] = ||w||  (if norm_idx >= 0)
   (three-term recurrence :245-247 with m <= 2, PRO projection :255-256 with m = j) */
int qsb_lz_axpy_norm(int dtype, int64_t n, int m, const void* X, int64_t ldx, void* w,
                     double* scal, int coef_idx, int norm_idx, int flags, void* scratch, void* stream);

/* dst <- src / Re(scal[idx])   (v_{j+1} = w / beta :268-270, _normalize_to :154) */
int qsb_lz_scale(int dtype, int64_t n, const void* src, void* dst, const double* scal, int idx,
                 void* stream);

/* dst <- sum_{i<m} y_host[i] * X_i ; scal[norm_idx] = ||dst||   (Ritz vector :293-294) */
int qsb_lz_combine(int dtype, int64_t n, int m, const void* X, int64_t ldx, const double* y_host,
                   void* dst, double* scal, int norm_idx, int flags, void* scratch, void* stream);

/* ---- utilities ---------------------------------------------------------- */
/* dst (contiguous, shape dims[0..2]) <- src with element strides; optional conj */
int qsb_strided_copy3(int dtype, const int64_t dims[3], const void* src, const int64_t src_stride[3],
                      void* dst, int conj, void* stream);

/* Plain batched GEMM on the library's FP64 tensor-core core, exposed for tests and
   the peak probe:  C[z][m][n] = sum_p sum_k opA(A[z,p][m][k]) * opB(B[z,p][k][n]).
   Element strides; for each of A and B one of the two strides must be 1. */
typedef struct {
    int dtype;
    int M, N, K, P, Z;
    const void* A; int64_t a_ms, a_ks, a_ps, a_zs; int conjA;
    const void* B; int64_t b_ns, b_ks, b_ps, b_zs; int conjB;
    void*       C; int64_t c_ms, c_ns, c_zs;
    int force_path;         /* 0 = auto, 1 = TMA (error if not eligible), 2 = cp.async */
    int accumulate;         /* 1: C += A.B */
    /* Optional real scale factors fused into the epilogue (NULL: none; not with accumulate):
       C[z][m][n] = row_scale[m / row_div] * col_scale[n % col_mod] * (A.B)[z][m][n].
       This is the Schmidt-weight scaling of the two-site wavefunction build (dmrg.py:633-660): theta's rows (l, s) are
       scaled with sWeight[l] in a right sweep (row_div = d), its columns (t, r) with sWeight[r] in a left sweep
       (col_mod = chi_r) -- no separate pass over a site tensor. */
    const double* row_scale; int row_div;
    const double* col_scale; int col_mod;
} qsb_gemm_desc;
int qsb_gemm(const qsb_gemm_desc* g, void* stream);

/* sizeof of the descriptor structs as this library was compiled: a binding asserts them against its own mirror */
size_t qsb_sizeof_heff_desc(void);
size_t qsb_sizeof_env_desc(void);
size_t qsb_sizeof_gemm_desc(void);

/* Register-resident DMMA issue-rate probe: runs `iters` rounds of independent
   mma.sync.m8n8k4.f64 on every SM and returns the flops issued in *flops.
   Timed by the caller with CUDA events -> the FP64 tensor-pipe peak of this GPU. */
int qsb_probe_dmma(int iters, double* flops, void* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QSB200_H */

/*
 * caldera_b200.h -- C ABI of libcaldera_b200.so (sm_100a CUDA kernels for the CALDERA
 * per-layer decomposition hot path).
 *
 * Every entry point takes raw DEVICE pointers, plain sizes and a cudaStream_t passed as
 * void*.  Nothing here allocates persistent device memory: outputs and scratch are owned
 * by the caller (size from the *_workspace_bytes() queries).  All calls are asynchronous
 * on `stream` unless stated otherwise.  Return value: 0 = ok, <0 = bad argument (the
 * Python shim raises the reference's exception class before launching), >0 = CUDA runtime
 * error (1000 + cudaError_t) or numerical failure.
 *
 * Reference interfaces replaced (RCR = /root/reference/rank-constrained-regression-main/src):
 *   cb_quantize_f32        LowMemoryQuantizer.quantize_block, uniform branch
 *                          RCR/caldera/utils/quantization.py:244-268 (+ :93-101)
 *   cb_dequantize_f32      LowMemoryQuantizer.dequantize_block, uniform branch
 *                          RCR/caldera/utils/quantization.py:290-295, 103-105, 306-307
 *   cb_pack_codes / cb_unpack_codes
 *                          new packed format; bit order of quantization.py:152, 217-220
 *   cb_caldera_layer       caldera()                 RCR/caldera/decomposition/alg.py:24-112
 *                          (maybe_update_Q :253-283, update_LR :128-198, LR_init :201-235,
 *                           quantize_matrix :245-250, activation_aware_error :286-302)
 *   cb_lowrank_init        LR_init                   RCR/caldera/decomposition/alg.py:201-235
 *   cb_lplr_iter           one pass of the LPLR loop RCR/caldera/decomposition/alg.py:160-188
 *   cb_weighted_error      activation_aware_error    RCR/caldera/decomposition/alg.py:286-302
 *   cb_convex_prox_iters   solve_convex_optimization  RCR/convex_caldera/decomposition/convex_caldera.py:128-241
 *   cb_quantize_residual_f32  quantize_residual        RCR/convex_caldera/decomposition/convex_caldera.py:342-373
 *   cb_sum_stats           kappa / c / certificates     RCR/convex_caldera/decomposition/convex_caldera.py:120-123, 404-406
 */
#ifndef CALDERA_B200_H
#define CALDERA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_VERSION 100

/* status codes */
#define CB_OK 0
#define CB_ERR_ARG (-1)          /* null pointer, negative size, bad enum           */
#define CB_ERR_BITS (-2)         /* bits not in {2,4,8,16}            -> AssertionError */
#define CB_ERR_BLOCK (-3)        /* numel % block != 0                -> ValueError     */
#define CB_ERR_WORKSPACE (-4)    /* workspace too small                                 */
#define CB_ERR_UNSUPPORTED (-5)  /* valid in the reference, not built here              */
#define CB_ERR_CUDA_BASE 1000    /* 1000 + cudaError_t                                  */
#define CB_ERR_NUMERIC 2000      /* e.g. Cholesky breakdown that the ridge retry could not fix */

int cb_version(void);
/* Human-readable text for a status code (static storage). */
const char* cb_status_string(int status);
/* Number of kernels this library has launched in this process (all streams). */
int64_t cb_kernel_launch_count(void);
/* Adds n to that counter: a caller replaying a captured CUDA graph of library kernels reports the
 * kernels the replay launched (the library cannot see graph replays). */
void cb_note_launches(int64_t n);

/* ------------------------------------------------------------------------- quantiser */

/* Block-wise abs-max uniform quantiser.
 *   x        rows x cols fp32, element (r,c) at x[r*stride_r + c*stride_c] (any strides;
 *            the reference quantises the non-contiguous view L.T, alg.py:171)
 *   bits     2, 4, 8 or 16;  levels = 2^(bits-1) - 1
 *   block    elements per block in row-major flattened order; 0 = one block (whole tensor)
 *   eps      scale floor (reference: 1e-8)
 *   codes    optional out, int8 (bits<=8) or int16 (bits==16), numel entries, row-major
 *   packed   optional out, cb_packed_bytes(numel,bits) bytes, offset-binary MSB-first
 *   scales   out, numel/block fp32
 *   dequant  optional out, numel fp32 row-major: (code/levels)*scale
 * Per element: s = max(absmax(block), eps); code = rint((x / s) * levels)  (IEEE divide,
 * IEEE multiply, round-half-even) -- bit-exact with the reference.
 * With only `packed` and `scales` requested (codes == NULL, dequant == NULL) the call is the quantise + pack pass of
 * the wire format: one read of x, bits/8 + 4/block bytes written per element.  A contiguous 2-bit tensor of at least a
 * few million elements is then staged through shared memory by bulk asynchronous copies and the launch carries the
 * programmatic-stream-serialisation attribute (the kernel waits for its predecessor in the stream before it touches
 * global memory, so ordering on `stream` is unchanged).
 */
int cb_quantize_f32(const float* x, int64_t rows, int64_t cols, int64_t stride_r, int64_t stride_c,
                    int bits, int64_t block, float eps,
                    void* codes, uint8_t* packed, float* scales, float* dequant, void* stream);

/* Inverse: out[i] = (code[i] / levels) * scales[i / block].  Exactly one of codes/packed
 * must be non-null. */
int cb_dequantize_f32(const void* codes, const uint8_t* packed, const float* scales,
                      int64_t numel, int bits, int64_t block, float* out, void* stream);

/* NormalFloat codebooks, LowMemoryQuantizer(method="nf4" / "nf2") (RCR/caldera/utils/quantization.py:32-90,
 * 270-279, 296-298): per block of `block` consecutive elements s = max(absmax, eps); idx = number of
 * thresholds (host array, midpoints of neighbouring levels) that x / s exceeds; dequant = levels[idx] * s.
 * x contiguous (the reference flattens), idx uint8, scales fp32 (numel / block). */
int cb_quantize_nf_f32(const float* x, int64_t numel, int64_t block, const float* thresholds_host, int n_thresholds,
                       float eps, uint8_t* idx, float* scales, void* stream);
int cb_dequantize_nf_f32(const uint8_t* idx, const float* scales, int64_t numel, int64_t block,
                         const float* levels_host, int n_levels, float* out, void* stream);

size_t cb_packed_bytes(int64_t numel, int bits);
int cb_pack_codes(const void* codes, int64_t numel, int bits, uint8_t* packed, void* stream);
int cb_unpack_codes(const uint8_t* packed, int64_t numel, int bits, void* codes, void* stream);

/* ------------------------------------------------------------------------- stages */

/* kind of the Hessian argument */
#define CB_H_IDENTITY 0   /* h == NULL                                   */
#define CB_H_DIAG 1       /* h = n fp32 diagonal entries                 */
#define CB_H_DENSE 2      /* h = n x n fp32 row-major symmetric matrix   */

/* Scans an n x n matrix: *is_diag (device int) = 1 when every off-diagonal entry of
 * (H + H^T)/2 is exactly zero; diag[n] receives the diagonal.  Lets callers that pass
 * torch.diag_embed(h) (main.py:163-165) take the diagonal fast path. */
int cb_hessian_probe(const float* H, int64_t n, float* diag, int* is_diag, void* stream);

/* num = sum_ij w_j * (W - Q - L R)_ij^2 and den = sum_ij w_j * W_ij^2 accumulated into
 * the device doubles out_num / out_den (caller zeroes them).  Q given as codes + scale
 * (or NULL for Q = 0); L (m x r), R (r x n) may be NULL for LR = 0.  For CB_H_DENSE the
 * weights are the full quadratic form tr(E H E^T). */
int cb_weighted_error(const float* W, int64_t m, int64_t n,
                      const void* q_codes, int q_bits, const float* q_scale,
                      const float* L, const float* R, int64_t r,
                      const float* h, int h_kind,
                      double* out_num, double* out_den, void* ws, size_t ws_bytes, void* stream);
size_t cb_weighted_error_workspace_bytes(int64_t m, int64_t n, int64_t r, int h_kind);

/* Rank-r factors minimising ||(A - L R) H^(1/2)||_F (LR_init, alg.py:201-235) by
 * randomized subspace iteration with oversampled width q_width, `niter` power iterations,
 * CholeskyQR re-orthonormalisation and a Rayleigh-Ritz step (one-sided Jacobi).
 *   aware != 0: L has orthonormal columns, R = L^T A           (== U_r, S_r V_r^T H^(-1/2))
 *   aware == 0: L = U_r sqrt(S_r), R = sqrt(S_r) V_r^T of A itself (H ignored)
 *   warm: optional q_width x m fp32 basis from a previous call (in/out), NULL = Gaussian start
 */
int cb_lowrank_init(const float* A, int64_t m, int64_t n, const float* h, int h_kind,
                    int64_t r, int64_t q_width, int niter, uint64_t seed, int aware,
                    float* L, float* R, float* sigma /* r, optional */,
                    void* ws, size_t ws_bytes, void* stream);
size_t cb_lowrank_init_workspace_bytes(int64_t m, int64_t n, int64_t r, int64_t q_width, int h_kind);

/* One iteration of the LPLR loop of update_LR (alg.py:160-188) from an injected state -- the stage the
 * survey's test pyramid asks for.  Given the residual res = W - Q (m x n), the Hessian weights and the
 * current R (r x n):
 *   L_pre  = argmin_L ||(res - L R) H^(1/2)||_F   (alg.py:163; normal equations + Cholesky here)
 *   L_idxs, L_scale, L_hat = whole-tensor quantisation of L_pre^T (alg.py:171-172; L_idxs has r*m codes in the
 *                            order of (L^T).flatten(), int8 for l_bits <= 8 else int16; L_hat is m x r)
 *   R_pre  = argmin_R ||res - L_hat R||_F         (alg.py:175)
 *   R_idxs, R_scale, R_hat = quantisation of R_pre (alg.py:179-180)
 *   err_sq = ||(res - L_hat R_hat) H_sqrt||_F^2   (alg.py:182; H_sqrt := H when aware == 0, alg.py:50)
 * Every output pointer may be NULL.  status (3 device ints, optional): Cholesky ridge retries, unused, tcgen05
 * watchdog.  h_kind: CB_H_IDENTITY or CB_H_DIAG. */
size_t cb_lplr_iter_workspace_bytes(int64_t m, int64_t n, int64_t r, int l_bits, int r_bits, int use_tensor_cores);
int cb_lplr_iter(const float* res, int64_t m, int64_t n, const float* h, int h_kind, int aware, int64_t r,
                 int l_bits, int r_bits, const float* R_in, float* L_pre, void* L_idxs, float* L_scale, float* L_hat,
                 float* R_pre, void* R_idxs, float* R_scale, float* R_hat, double* err_sq, int use_tensor_cores,
                 int* status, void* ws, size_t ws_bytes, void* stream);

/* Small dense helpers (exported for tests and for callers that build their own loops). */
/* G (q x q, symmetric, row-major, destroyed) -> Lc in the lower triangle of G and
 * Linv = Lc^-1 (q x q lower triangular, may be NULL).  On breakdown the factorisation is
 * retried on G + ridge*mean(diag)*I with ridge = 1e-6, 1e-4, 1e-2 and *status gets the
 * number of retries (device int, may be NULL). */
int cb_cholesky_inverse_f32(float* G, int64_t q, float* Linv, int* status, void* stream);
/* Eigen-decomposition of the SPD matrix G = Lc Lc^T from its Cholesky factor by one-sided
 * Jacobi on Lc's columns.  evals[q] descending, evecs row k = k-th eigenvector.
 * work: q*q + q + 8 floats of scratch (column storage, norms, sweep counters). */
int cb_jacobi_eigh_from_chol_f32(const float* Lc, int64_t q, float* evals, float* evecs,
                                 float* work, int* sweeps, void* stream);
/* The sigma_reg shift of alg.py:57-64 for a dense symmetric H (n x n, row-major, in place):
 * H += max(0, sigma_reg - lambda_min(H)) I.  lambda_min comes from <= 96 Lanczos steps with full reorthogonalisation
 * and bisection on the tridiagonal matrix (the reference reads it off a full eigh).  stats (2 device floats, optional):
 * the shift applied and the lambda_min estimate. */
size_t cb_min_eig_shift_workspace_bytes(int64_t n);
int cb_min_eig_shift_f32(float* H, int64_t n, float sigma_reg, float* stats, void* ws, size_t ws_bytes, void* stream);
/* C = alpha * op(A) op(B) (+ C) with arbitrary element strides: A(i,k) = A[i*a_rs + k*a_cs],
 * B(k,j) = B[k*b_rs + j*b_cs], C(i,j) = C[i*c_rs + j*c_cs].  fp32 SIMT kernel: the path for
 * shapes the tcgen05 tiles do not cover and the in-library reference for them. */
int cb_sgemm_strided(int64_t M, int64_t N, int64_t K, float alpha,
                     const float* A, int64_t a_rs, int64_t a_cs,
                     const float* B, int64_t b_rs, int64_t b_cs,
                     float* C, int64_t c_rs, int64_t c_cs, int accumulate, void* stream);

/* C (M x N fp32, ldc) = alpha * A (M x K) * B (N x K)^T with bf16 operands, both K-major
 * (lda, ldb in elements, multiples of 8), fp32 accumulation in tensor memory: the
 * tcgen05 / TMEM / TMA contraction kernel (csrc/gemm_tc.cu).  splitk <= 0 picks the K split
 * that fills the machine.  Split-K is atomics-free (partial tiles meet in `workspace`, device
 * memory of cb_gemm_bf16_tn_workspace_bytes() bytes, and are summed in slice order), so equal
 * inputs give equal bits; with workspace == NULL the contraction runs unsplit (splitk > 1 is then
 * CB_ERR_WORKSPACE).  *error_flag (device int, may be NULL) is set if the in-kernel pipeline
 * watchdog fired.  probe_flags: 0; 1 = measurement aid, run the TMA pipeline without issuing MMAs
 * (C undefined).  Exported for validation against cb_sgemm_strided. */
size_t cb_gemm_bf16_tn_workspace_bytes(void);
int cb_gemm_bf16_tn(int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                    const void* B_bf16, int64_t ldb, float* C, int64_t ldc, int splitk, int probe_flags,
                    int* error_flag, void* workspace, size_t workspace_bytes, void* stream);
/* The same contraction with the bf16 epilogues the layer driver uses: Cb (M x N row-major bf16, ldcb) and/or
 * Ct (N x M bf16, the transpose, ldct), either may be NULL; optional per-column / per-row fp32 scaling
 * (NULL = none).  When the outputs are vector-aligned (N % 8 == 0, ldcb % 8 == 0; M % 4 == 0, ldct % 4 == 0)
 * the tile is staged through shared memory and written as whole rows.  exec_mode (CB_MODE_*): the grid policy of
 * the call, as in cb_caldera_params.  Exported for validation. */
int cb_gemm_bf16_tn_bf16out(int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                            const void* B_bf16, int64_t ldb, void* Cb_bf16, int64_t ldcb, void* Ct_bf16,
                            int64_t ldct, const float* colscale, const float* rowscale, int exec_mode, int* error_flag,
                            void* stream);
/* The batched, persistent CTA-pair form of the same contraction (csrc/gemm_tc2.cu): for b < batch,
 *   C[b] (M x N) = alpha * A[b] (M x K) * B[b] (N x K)^T  [* rowscale[b][i] * colscale[b][j]]
 * with `tcgen05.mma.cta_group::2` on 256 x N_TILE tiles (N_TILE = min(256, N rounded up to 16)), two accumulators
 * in tensor memory so that the epilogue of a tile overlaps the main loop of the next, and the batch as a third
 * tensor-map dimension.  stride_*_bytes: distance between two batch items of each array (multiples of 16 for A and
 * B).  Any subset of C (fp32), Cb (bf16 row-major), Ct (bf16, N x M) may be given.  max_clusters <= 0: one cluster
 * per SM pair.  tile_counter: two zeroed device ints for the dynamic tile scheduler (the kernel leaves them zero, so
 * launches that follow each other on one stream can share them), or NULL for a fixed tile -> cluster assignment; the
 * results do not depend on it.  This is what the batched layer driver runs; exported for validation. */
int cb_gemm_bf16_tn_batched(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                            int64_t stride_a_bytes, const void* B_bf16, int64_t ldb, int64_t stride_b_bytes, float* C,
                            int64_t ldc, int64_t stride_c_bytes, void* Cb_bf16, int64_t ldcb, int64_t stride_cb_bytes,
                            void* Ct_bf16, int64_t ldct, int64_t stride_ct_bytes, const float* colscale,
                            int64_t stride_col_bytes, const float* rowscale, int64_t stride_row_bytes, int max_clusters,
                            int* tile_counter, int* error_flag, void* stream);
/* ---- consumer of the packed decomposition (the reference reconstructs a dense matrix,
 * main.py:197 / README.md:182 `W_hat = Q + L @ R`, and multiplies with that) ----
 * y[T, m] = global_scale * x[T, n] * (Q + L R)^T with Q = (codes / levels) * q_scale read from its packed
 * form (q_bits in {2, 4, 8}, MSB-first as written by cb_quantize_f32 / cb_caldera_layer; one scale per
 * tensor, device pointer), L (m x r) and R (r x n) fp32 row-major (r may be 0).  One tcgen05 kernel expands
 * the codes tile by tile in shared memory (the dense Q never exists in HBM) and runs the rank-r part as extra
 * K blocks into a second TMEM accumulator; operands are rounded to bf16 (codes are exact), accumulation is
 * fp32.  Requires n % 64 == 0 and r % 8 == 0.  x, y row-major fp32. */
size_t cb_packed_linear_workspace_bytes(int64_t T, int64_t m, int64_t n, int64_t r);
int cb_packed_linear_f32(const float* x, int64_t T, int64_t n, const uint8_t* q_packed, int q_bits,
                         const float* q_scale, const float* L, const float* R, int64_t m, int64_t r,
                         float global_scale, float* y, int* error_flag, void* ws, size_t ws_bytes, void* stream);
/* ---- Hessian accumulation (the reference does `a_aT = activations @ activations.T` in fp64
 * on the CPU for every calibration sample, main.py:307-311, and convex_caldera.py:108) ----
 * H (n x n fp32, row-major) += X^T X and/or hdiag (n fp32) += sum_t X[t, j]^2 for a batch X (T x n fp32,
 * row-major) of activations; either output may be NULL.  The product runs on the tcgen05 kernel with
 * split-bf16 operands (relative error ~1e-5 of an fp64 product).  The caller zeroes H / hdiag before the
 * first batch and divides by the sample count at the end. */
size_t cb_hessian_accumulate_workspace_bytes(int64_t T, int64_t n);
int cb_hessian_accumulate_f32(const float* X, int64_t T, int64_t n, float* H, float* hdiag, int* error_flag,
                              void* ws, size_t ws_bytes, void* stream);
/* ---- Hadamard pre-rotation (main.py:79-133 builds dense normalised Hadamard matrices with
 * scipy and multiplies on the CPU in fp64) ----
 * out (prows x pcols fp32) = H1 * pad(W) * H2 with H_k the normalised Sylvester-Hadamard matrix of order
 * prows / pcols (powers of two >= rows / cols, <= 32768) and pad() the zero padding of main.py:84-90, as a
 * fast Walsh-Hadamard transform.  The transform is its own inverse (main.py:124-129): apply it to the padded
 * matrix again and keep the leading rows x cols block. */
size_t cb_hadamard_workspace_bytes(int64_t prows, int64_t pcols);
int cb_hadamard_transform_f32(const float* W, int64_t rows, int64_t cols, float* out, int64_t prows, int64_t pcols,
                              void* ws, size_t ws_bytes, void* stream);
/* How one layer uses the machine: the `exec_mode` field of cb_caldera_params (part of the call, never process-wide
 * state).  Results are bitwise reproducible within a mode and agree to rounding level between modes (different
 * K-split counts and eigensolver sweep order).
 *   CB_MODE_LATENCY (0): one layer at a time should finish as early as possible -- contractions spread over ~all
 *     SMs, the 8-CTA cluster eigensolver.
 *   CB_MODE_THROUGHPUT (1): many independent layers are in flight on different streams -- contractions use ~32-CTA
 *     grids with the widest tiles (3x less SM time and fabric traffic per flop, four of them fit side by side), the
 *     single-CTA eigensolver (3x less SM time).
 * The batched driver (cb_caldera_batch) has one policy and ignores the field. */
#define CB_MODE_LATENCY 0
#define CB_MODE_THROUGHPUT 1
/* fp32 (rows x cols, ldx) -> bf16 copy Y (ldy) and/or transposed copy Yt (cols x rows, ldyt),
 * optionally scaling column c by colscale[c] first. */
int cb_convert_bf16(const float* X, int64_t rows, int64_t cols, int64_t ldx, void* Y_bf16, int64_t ldy,
                    void* Yt_bf16, int64_t ldyt, const float* colscale, void* stream);

/* ------------------------------------------------------------------------- fused driver */

typedef struct cb_caldera_params {
  int32_t compute_q;        /* CalderaParams.compute_quantized_component */
  int32_t compute_lr;       /* CalderaParams.compute_low_rank_factors    */
  int32_t q_bits, l_bits, r_bits;
  int32_t rank;
  int32_t iters;
  int32_t lplr_iters;
  int32_t aware;            /* activation_aware_LR                        */
  int32_t n_order;          /* len(update_order), <= 8                    */
  int32_t order[8];         /* 0 = "Q", 1 = "LR"                          */
  int32_t rand_svd;         /* 1: q=2r, niter=2 like torch.svd_lowrank; 0: accurate mode */
  float sigma_reg;
  int32_t scale_w;          /* caldera(..., scale_W=)                     */
  float global_scale_in;    /* >0: use this instead of computing sqrt(mean(W^2)) */
  int64_t q_block;          /* 0 = whole-tensor scale (reference semantics, alg.py:247);
                               >0 = per-block scales (opt-in extension)   */
  int32_t sketch_width;     /* 0 = default (2*rank)                       */
  int32_t power_iters;      /* of a cold (random-start) rank-r step; <0 = default (2 if rand_svd else 12) */
  int32_t power_iters_warm; /* of a step warm-started from the previous outer iteration's basis;
                               <0 = default (power_iters if that is given, else 2 if rand_svd else 3) */
  int32_t warm_start;       /* reuse the previous outer iteration's basis */
  int32_t use_tensor_cores; /* 1: bf16 tcgen05 contractions where the shape allows (dims % 8 == 0,
                               m, n >= 256); 0: fp32 SIMT contractions everywhere        */
  int32_t exec_mode;        /* CB_MODE_LATENCY (0) or CB_MODE_THROUGHPUT (1): grid policy of this call     */
  uint64_t seed;
} cb_caldera_params;

typedef struct cb_caldera_out {
  /* all device pointers, caller-allocated */
  float* Q;            /* m x n fp32, best iterate (scaled space)                  */
  float* L;            /* m x r                                                     */
  float* R;            /* r x n                                                     */
  void* Q_idxs;        /* int8/int16 m*n codes of the best iterate                  */
  float* Q_scale;      /* numel/q_block (or 1) scales                               */
  uint8_t* Q_packed;   /* optional, cb_packed_bytes(m*n, q_bits)                    */
  void* L_idxs;        /* int8/int16, r*m, order of (L^T).flatten(); only if l_bits<16 or r_bits<16 */
  void* R_idxs;        /* int8/int16, r*n                                           */
  float* L_scale;      /* 1 */
  float* R_scale;      /* 1 */
  uint8_t* L_packed;   /* optional */
  uint8_t* R_packed;   /* optional */
  float* W_scaled;     /* optional m x n: W / global_scale (CalderaDecomposition.W) */
  float* errors;       /* iters * n_order floats, in sub-step order                 */
  const uint64_t* seed_dev; /* optional device uint64 added to params.seed: lets a captured CUDA graph of the
                          layer be replayed with a different seed per layer                       */
  float* scalars;      /* 8 floats: [0]=global_scale [1]=min_error [2]=best_step
                          [5],[6],[7] hold int32 bit patterns: cholesky ridge retries (max over
                          the layer), jacobi sweeps (last solve), tcgen05 pipeline watchdog */
} cb_caldera_out;

size_t cb_caldera_layer_workspace_bytes(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind);
/* Whole per-layer decomposition, enqueued on `stream` with no host synchronisation. */
int cb_caldera_layer(const cb_caldera_params* p, const float* W, int64_t m, int64_t n,
                     const float* h, int h_kind, const cb_caldera_out* out,
                     void* ws, size_t ws_bytes, void* stream);

/* The same decomposition for `batch` same-shape layers advancing in lock step on one stream (the model-level job:
 * a 7B model has 128 + 64 + 32 layers of three shapes).  Layer b reads W + b * stride_bytes and h + b * stride_bytes,
 * writes through every pointer of `out` + b * stride_bytes and uses the workspace ws + b * stride_bytes
 * (cb_caldera_layer_workspace_bytes() each): one slab per layer, stride_bytes a multiple of 256.  Each contraction of
 * the rank-r step, each small factorisation and each bookkeeping step is then ONE launch for the whole batch (the
 * batched CTA-pair tcgen05 contraction; one CTA per layer for Cholesky / Jacobi).  A layer's result does not depend
 * on the batch size or on its position in the batch.  Supported (cb_caldera_batch_supported() != 0): tensor-core
 * path, Q and LR both computed, identity / diagonal Hessian, rank <= 192; other configurations return
 * CB_ERR_UNSUPPORTED and go through cb_caldera_layer. */
int cb_caldera_batch_supported(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind);
int cb_caldera_batch(const cb_caldera_params* p, int batch, int64_t stride_bytes, const float* W, int64_t m, int64_t n,
                     const float* h, int h_kind, const cb_caldera_out* out, void* ws, size_t ws_bytes, void* stream);

/* bitsandbytes-style asymmetric block quantisers `bbint4` / `bbint2` of the QuantizerFactory surface
 * (RCR/caldera/utils/quantization.py:107-243): per block of `block` consecutive elements mean / unbiased std,
 * outliers (|x - mean| > 6 max(std, eps)) go to a side table and are replaced by the mean, then
 * code = clamp(rint((x - min) / scale), 0, 2^bits - 1) with scale = max((max - min) / (2^bits - 1), eps), packed
 * MSB-first (:152, :217-220).  cb_quantize_bbint_f32 writes the packed codes (numel * bits / 8 bytes), block_min and
 * scales (numel / block floats), the outlier count of every block (counts) and its exclusive scan (offsets,
 * numel / block + 1 entries, the last one the total).  cb_bbint_outliers_f32 then fills the side table (values and
 * (block row, column) pairs in row-major order, the order of torch.nonzero) sized from that total.
 * cb_dequantize_bbint_f32: code * scale + min, outliers restored (:156-173, :223-243).  The reference's CSV log of
 * outlier counts (:126-137) is not written. */
int cb_quantize_bbint_f32(const float* x, int64_t numel, int64_t block, int bits, float eps, uint8_t* packed,
                          float* block_min, float* scales, int* counts, int64_t* offsets, void* stream);
int cb_bbint_outliers_f32(const float* x, int64_t numel, int64_t block, float eps, const int* counts,
                          const int64_t* offsets, float* values, int64_t* indices, void* stream);
int cb_dequantize_bbint_f32(const uint8_t* packed, const float* block_min, const float* scales, int64_t numel,
                            int64_t block, int bits, const float* outlier_values, const int64_t* outlier_indices,
                            int64_t n_outliers, float* out, void* stream);

/* ------------------------------------------------------------------------- Convex-CALDERA */

/* out3[0] += sum x, out3[1] += sum x^2, out3[2] += sum (x - y)^2 (y may be NULL); device doubles the
 * caller zeroes.  kappa = ||W||_F, c = 0.1 var(W) (convex_caldera.py:120-123) and the
 * certificates (convex_caldera.py:404-406) are built from these. */
int cb_sum_stats(const float* x, const float* y, int64_t numel, double* out3, void* stream);

/* n_iters accelerated proximal-gradient iterations of the reduced Convex-CALDERA program
 *   min 1/2 ||(W-L-R) diag(h)^(1/2)||_F^2 + mu ||L||_* + lambda max(q0, ||R||_F^2 / kappa)
 * (mu < 0 selects the constrained form ||L||_* <= tau_star).  Replaces the CVXPY/SCS solve of
 * solve_convex_optimization (convex_caldera.py:128-241); see oracle/convex_oracle.py for the
 * reduction.  State L, Lp, R, Rp (m x n, zero-initialised by the caller for a cold start) is
 * updated in place; Lf (m x rank_cap) and Rf (rank_cap x n) receive U sqrt(S), sqrt(S) V^T of
 * the last thresholding argument, svals[0:rank_cap] the thresholded singular values of L and
 * svals[rank_cap:2*rank_cap] the ratios s'/S.  scalars (device doubles, >= 5): [0] ||L||_*,
 * [1] radial factor, [2] ||R||_F^2, [3] smooth term, [4] scratch.  *theta_io is the host-side
 * FISTA momentum state (start at 1.0; reset to 1.0 to restart). */
int cb_convex_prox_iters(const float* W, const float* h, int64_t m, int64_t n, float mu, float tau_star,
                         float lambda_reg, float kappa, float q0, float step_t, int64_t rank_cap,
                         int64_t q_width, int power_iters, uint64_t seed, int warm, int use_tensor_cores,
                         int n_iters, double* theta_io, float* L, float* Lp, float* R, float* Rp,
                         float* Lf, float* Rf, float* svals, double* scalars, void* ws, size_t ws_bytes,
                         void* stream);
size_t cb_convex_prox_workspace_bytes(int64_t m, int64_t n, int64_t rank_cap, int64_t q_width,
                                      int use_tensor_cores);

/* Dense (non-diagonal) Hessians, e.g. the Gram matrix X^T X of calibration data (convex_caldera.py:103-117).
 * cb_convex_dense_prepare: Hs = (H + H^T) / 2 (:111) shifted by max(0, floor - lambda_min) on its diagonal (the
 * reference clamps the eigenvalues at 1e-8, :113; for a positive semi-definite H the two differ by at most `floor`
 * on every eigenvalue); stats3 (device floats) = {shift, lambda_min, lambda_max}, the extreme eigenvalues from a
 * Lanczos run (<= 96 steps, full reorthogonalisation).  ws: cb_min_eig_shift_workspace_bytes(n).
 * cb_convex_prox_iters_dense: cb_convex_prox_iters with the smooth term 1/2 ||(W-L-R) Hs^(1/2)||_F^2 for the dense
 * symmetric Hs; no square root is formed: the gradient is (L+R-W) Hs (one fp32 contraction per iteration) and
 * scalars[3] (1/2 sum (E Hs) (.) E) is evaluated for the last iterate of the call.  step_t <= 1 / (2 lambda_max). */
int cb_convex_dense_prepare(const float* H, int64_t n, float floor, float* Hs, float* stats3, void* ws, size_t ws_bytes,
                            void* stream);
int cb_convex_prox_iters_dense(const float* W, const float* Hs, int64_t m, int64_t n, float mu, float tau_star,
                               float lambda_reg, float kappa, float q0, float step_t, int64_t rank_cap,
                               int64_t q_width, int power_iters, uint64_t seed, int warm, int use_tensor_cores,
                               int n_iters, double* theta_io, float* L, float* Lp, float* R, float* Rp,
                               float* Lf, float* Rf, float* svals, double* scalars, void* ws, size_t ws_bytes,
                               void* stream);

/* quantize_residual (convex_caldera.py:342-373): delta = 2 max|R| / (2^bits - 1) (max|R| / 2^15 for
 * bits == 16), R_int = clamp(rint(R / delta), +-(2^(bits-1) - 1)), Rq = delta * R_int and, when
 * base/Wc are given, Wc = base + Rq (the reconstruction L* + delta R_int, :484-485).
 * scratch: one device float. */
int cb_quantize_residual_f32(const float* R, const float* base, int64_t rows, int64_t cols, int bits,
                             float* Rq, float* Wc, float* delta_out, float* scratch, void* stream);

/* out = X scaled along rows (axis 0) or columns (axis 1) by v: mode 0 multiply, 1 divide. */
int cb_scale_f32(const float* X, int64_t rows, int64_t cols, const float* v, int axis, int mode, float* out,
                 void* stream);

/* ------------------------------------------------------------------------- measurement aids
 * Compiled only into libcaldera_b200_measure.so (`python -m ee274_convexcaldera_llm_quantization_b200.build --measure`,
 * -DCB_MEASURE) for scripts/probe_*.py; the release library exports none of them and has no process-wide setter.  That
 * build also honours CB_DEBUG_SKIP (drops whole kernel classes for knock-out timing; results are then garbage). */
#ifdef CB_MEASURE
/* 64-wide K blocks fetched per TMA instruction in the narrow-tile contractions (1 or 2, default 2). */
void cb_set_gemm_kblocks(int n);
/* 0: direct-store epilogue instead of the shared-memory staged one (same results). */
void cb_set_gemm_staged_epilogue(int on);
/* When non-NULL, CTA (0,0,0) of every tcgen05 contraction writes clock64 stamps to this device buffer of 12 int64
 * ([0..7]: entry, prologue done, last load issued, first stage landed, last stage landed, accumulator complete,
 * epilogue stores issued, exit; [8..9]: staged epilogue -- tile in shared memory, barrier passed). */
void cb_set_gemm_timing(void* stamps_dev);
/* When non-NULL the Cholesky kernel writes clock64 stamps (start, factor done, diagonal block inverses done, inverse
 * done; then 5 per panel for the first 3 panels) to this device buffer of 24 int64. */
void cb_set_chol_timing(void* stamps_dev);
/* Cycles that n_mma back-to-back tcgen05.mma (128 x bn x 16, bf16, operands resident in shared memory) take on an SM,
 * on a grid of `grid` CTAs.  out_cycles: two device int64 ([0] issue + drain, [1] issue only). */
int cb_probe_mma_rate(int bn, int n_mma, int distinct_k, int grid, void* out_cycles, void* stream);
/* The error pass of the outer loop alone (stages.cu: err_accum), with every optional operand: num += sum_ij w_j (Ws - Q -
 * LR)_ij^2, *amax_next = max |Ws - LR|.  codes: int8 (bits <= 8) or int16; LR, w, amax_next may be NULL. */
int cb_probe_err_pass(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR, const float* w,
                      int64_t m, int64_t n, double* num, float* amax_next, void* stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* CALDERA_B200_H */

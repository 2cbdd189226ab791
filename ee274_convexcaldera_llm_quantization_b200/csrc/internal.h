// Internal (non-ABI) interfaces shared between the translation units of libcaldera_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cb {

// A batch of same-shape layers advancing in lock step: layer b's buffers (inputs, outputs, workspace) sit b * stride
// bytes after layer 0's.  Stage functions take layer-0 pointers; n == 1 is the single-layer case.
struct Bt {
  int n = 1;
  int64_t stride = 0;
};
template <typename T> inline T* at(T* p, const Bt& bt, int b) {
  return (p == nullptr || b == 0) ? p : reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + (uintptr_t)((int64_t)b * bt.stride));
}

// Scratch for split-K: the partial tiles of every K slice, summed in slice order by a reduce launch
// that follows the contraction on the same stream (no floating-point atomics).  Without it, or when
// it is too small, the contraction runs unsplit.
struct SplitWs {
  float* buf = nullptr;
  size_t bytes = 0;
};
constexpr size_t kSplitWsBytes = 12u << 20;


// Execution policy of the single-layer driver: grid-size target of the tcgen05 contractions (gemm_tc.cu) and the
// eigensolver variant (smalldense.cu).  It is part of the call (cb_caldera_params.exec_mode), never process-wide
// state: an entry point installs it for its own thread while it enqueues (PolicyScope) and restores it on return.
struct ExecPolicy {
  int target_ctas = 120;     // ~120: one layer fills the machine (latency); ~32: several layers' grids side by side
  int jacobi_single = 0;     // 1: single-CTA eigensolver (3x less SM time), 0: 8-CTA cluster kernel (lower latency)
};
extern thread_local ExecPolicy tl_policy;
struct PolicyScope {
  ExecPolicy saved;
  explicit PolicyScope(int exec_mode) : saved(tl_policy) {
    if (exec_mode == 1) { tl_policy.target_ctas = 32; tl_policy.jacobi_single = 1; }       // CB_MODE_THROUGHPUT
    else { tl_policy.target_ctas = 120; tl_policy.jacobi_single = 0; }                     // CB_MODE_LATENCY
  }
  ~PolicyScope() { tl_policy = saved; }
};

// sgemm.cu -- C(i,j) = alpha * sum_k A(i,k) B(k,j) [* colscale[j]] (+ C), arbitrary strides
int sgemm(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t a_rs, int64_t a_cs,
          const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t c_rs, int64_t c_cs,
          bool accumulate, const float* colscale, cudaStream_t st, const SplitWs* sw = nullptr, const Bt& bt = Bt());

// smalldense.cu
int cholesky_inverse(float* G, int q, float* Linv, int* status, cudaStream_t st,
                     __nv_bfloat16* Linv_bf16 = nullptr, const Bt& bt = Bt());
int jacobi_eigh_from_chol(const float* Lc, int q, float* evals, float* evecs, float* work, int* sweeps,
                          cudaStream_t st, const Bt& bt = Bt());

// Hs (n x n symmetric, in place) += max(0, sigma_reg - lambda_min(Hs)) I with lambda_min from a Lanczos run
// (alg.py:57-64); work: min_eig_shift_workspace_floats(n) floats; stats (optional, nstats <= 3 floats): shift,
// lambda_min, lambda_max (the extreme Ritz values of the same run)
size_t min_eig_shift_workspace_floats(int64_t n);
int min_eig_shift(float* Hs, int64_t n, float sigma_reg, float* work, float* stats, cudaStream_t st, int nstats = 2);

// stages.cu -- fused element-wise stages of the outer loop (all asynchronous on `st`)
struct HessianVecs {
  const float* h;        // column weights of the error metric (n) or nullptr for all ones
  const float* sqrt_h;   // sqrt(h) or nullptr
  const float* inv_sqrt_h;
  const float* w_inner;  // weights of the LPLR inner error: h (aware) or h^2 (not aware, alg.py:50)
};

int sumsq(const float* x, int64_t numel, double* out, cudaStream_t st);
int finalize_global_scale(const double* sumsq, int64_t numel, float gs_in, int scale_w, float* scalars,
                          cudaStream_t st, const Bt& bt = Bt());
// amax (optional, zeroed by the caller): receives max |W / gs|, the abs-max the first Q update needs
int scale_and_den(const float* W, float* Ws, int64_t m, int64_t n, const float* gs, const float* h,
                  double* den, cudaStream_t st, float* amax = nullptr);
int prep_hessian_diag(const float* h_in, int64_t n, float sigma_reg, int aware, float* h_eff, float* sqrt_h,
                      float* inv_sqrt_h, float* w_inner, float* scratch, cudaStream_t st, const Bt& bt = Bt());
int resid_absmax(const float* Ws, const float* LR, int64_t numel, float* amax, cudaStream_t st);
int quant_err(const float* Ws, const float* LR, const float* h, int64_t m, int64_t n, const float* amax,
              float eps, int bits, void* codes, float* qscale, double* num, cudaStream_t st);
int form_y(const float* Ws, const void* codes, int bits, const float* qscale, const float* sqrt_h,
           int64_t m, int64_t n, float* Y, float* RES, cudaStream_t st);
int err_accum(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR,
              const float* w, int64_t m, int64_t n, double* num, cudaStream_t st, float* amax_next = nullptr);
int select_outer(double* num, const double* den, float* errors, int step, float* scalars, int* flags,
                 int all_updated, cudaStream_t st, const Bt& bt = Bt());
int select_inner(double* num, float* scalars, int* flags, int first, int last, cudaStream_t st, const Bt& bt = Bt());
int copy_if(const int* flag, void* dst, const void* src, size_t bytes, cudaStream_t st);
// up to 8 (dst, src, bytes) segments copied by one launch when *flag != 0
struct CopySegments {
  void* dst[8];
  const void* src[8];
  size_t bytes[8];
  int count = 0;
  void add(void* d, const void* s_, size_t b) {
    if (b == 0 || count >= 8 || d == nullptr) return;
    dst[count] = d; src[count] = s_; bytes[count] = b; ++count;
  }
};
// flag == nullptr: unconditional; a segment with src == nullptr is zero-filled
int copy_if_multi(const int* flag, const CopySegments& segs, cudaStream_t st, const Bt& bt = Bt());
int fill_randn(float* p, int64_t count, uint64_t seed, const uint64_t* seed_dev, cudaStream_t st);
int scale_cols(const float* X, int64_t rows, int64_t cols, const float* v, int mode, float* out, cudaStream_t st);
int scale_rows(const float* X, int64_t rows, int64_t cols, const float* v, int mode, float* out, cudaStream_t st);
int transpose_codes(const void* src, int64_t rows, int64_t cols, int elem_bytes, void* dst, cudaStream_t st);
int form_e(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR, int64_t m, int64_t n,
           float* E, cudaStream_t st);
int dot_accum(const float* a, const float* b, int64_t numel, double* out, cudaStream_t st);
int symmetrize(const float* H, int64_t n, float* Hs, cudaStream_t st);
int form_y_bf16(const float* Ws, const void* codes, int bits, const float* qscale, const float* sqrt_h, int64_t m,
                int64_t n, __nv_bfloat16* Yb, __nv_bfloat16* Ytb, float* RES, cudaStream_t st);
// quant_err + form_y_bf16 in one pass (see stages.cu)
int quant_form_y_bf16(const float* Ws, const float* LR, const float* h_err, const float* sqrt_h, int64_t m, int64_t n,
                      const float* amax, float eps, int bits, void* codes, float* qscale, double* num,
                      __nv_bfloat16* Yb, __nv_bfloat16* Ytb, float* RES, cudaStream_t st);
// whole-tensor quantise + dequantise (quantize_matrix, alg.py:245-250) of one contiguous fp32 tensor per layer of a
// batch: codes (int8 / int16), the clamped scale, the dequantised values.  amax_scratch: one float per layer.
int quantize_whole_batched(const float* x, int64_t numel, int bits, void* codes, float* scale_out, float* deq,
                           float* amax_scratch, cudaStream_t st, const Bt& bt);
int cvx_point(const float* W, const float* L, const float* Lp, const float* R, const float* Rp, const float* h,
              int64_t m, int64_t n, float beta, float t, float* VL, float* VR, double* vr_sumsq, cudaStream_t st);
// dense-Hessian variants: the step is split around the contraction G = (W - Y_L - Y_R) H
int cvx_resid(const float* W, const float* L, const float* Lp, const float* R, const float* Rp, int64_t m, int64_t n,
              float beta, float* D, cudaStream_t st);
int cvx_point_dense(const float* L, const float* Lp, const float* R, const float* Rp, int64_t m, int64_t n, float beta,
                    float t, float* VL /* in: G, out: V_L */, float* VR, double* vr_sumsq, cudaStream_t st);
int cvx_finish_dense(const float* W, const float* Lnew, const float* VR, int64_t m, int64_t n, const double* sc,
                     float* Rnew, float* E /* optional: W - L_new - R_new */, cudaStream_t st);
int scale_double(double* x, double f, cudaStream_t st);
int cvx_shrink(const float* sigma2, int r, float thresh, float tau_star, int constrained, const double* vr_sumsq,
               float t_lambda, float kappa, float q0, float* colw, float* s_out, double* sc, cudaStream_t st);
int cvx_finish(const float* W, const float* Lnew, const float* VR, const float* h, int64_t m, int64_t n, const double* sc,
               float* Rnew, double* smooth, cudaStream_t st);

// gemm_tc.cu -- tcgen05 path: C[M,N] (+)= alpha * A[M,K] * B[N,K]^T, bf16 K-major operands
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb);
int gemm_tc(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A, int64_t lda,
            const __nv_bfloat16* B, int64_t ldb, float* C, int64_t ldc, __nv_bfloat16* Cb, int64_t ldcb,
            __nv_bfloat16* Ct, int64_t ldct, const float* colscale, const float* rowscale, int splitk,
            int* error_flag, int* splits_used, cudaStream_t st, const SplitWs* sw = nullptr,
            int probe_flags = 0 /* bit 0: loads only (measurement aid) */);
// gemm_tc2.cu -- batched, persistent CTA-pair variant: C[b] = alpha * A[b] * B[b]^T for b < batch, byte strides between
// batch items (s*); outputs as in gemm_tc (fp32 C, bf16 Cb, transposed bf16 Ct, any subset)
struct Gemm2Batch {
  int64_t batch = 1, M = 0, N = 0, K = 0;
  float alpha = 1.f;
  const __nv_bfloat16* A = nullptr; int64_t lda = 0, sA = 0;
  const __nv_bfloat16* B = nullptr; int64_t ldb = 0, sB = 0;
  float* C = nullptr; int64_t ldc = 0, sC = 0;
  __nv_bfloat16* Cb = nullptr; int64_t ldcb = 0, sCb = 0;
  __nv_bfloat16* Ct = nullptr; int64_t ldct = 0, sCt = 0;
  const float* colscale = nullptr; int64_t sCol = 0;
  const float* rowscale = nullptr; int64_t sRow = 0;
  __nv_bfloat16* Cs = nullptr; int64_t sCs = 0; int split_mode = 0;   // split-bf16 copy of C, see gemm_tc2.cu
  int max_clusters = 0;          // 0 = one cluster per SM pair
  int* error_flag = nullptr;
  int* tile_counter = nullptr;   // 2 zeroed device ints for the dynamic tile scheduler (left zero by the kernel); null = static
};
bool gemm_tc2_supported(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb);
int gemm_tc2(const Gemm2Batch& g, cudaStream_t st);
int to_bf16(const float* X, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* Y, int64_t ldy,
            __nv_bfloat16* Yt, int64_t ldyt, const float* colscale, cudaStream_t st, const Bt& bt = Bt());

// quant.cu (C ABI, reused internally)
}  // namespace cb

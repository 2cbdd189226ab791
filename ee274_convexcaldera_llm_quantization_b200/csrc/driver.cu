// Host-side orchestration of one layer's decomposition: enqueues every stage of
// caldera() (RCR/caldera/decomposition/alg.py:24-112) on one stream with no host
// synchronisation.  Best-iterate selection (alg.py:105-107) and the LPLR inner best
// (alg.py:184-188) are device-side flags consumed by predicated copies, and the error
// trajectory is written to a device array the caller reads once at the end.
#include <new>
#include "common.cuh"
#include "internal.h"

namespace cb {

// ---------------------------------------------------------------- workspace arena
struct Arena {
  uint8_t* base;
  size_t off;
  size_t cap;
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base == nullptr ? nullptr : reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

static inline int code_bytes(int bits) { return bits <= 8 ? 1 : 2; }
static inline bool bits_ok(int b) { return b == 2 || b == 4 || b == 8 || b == 16; }

// ---------------------------------------------------------------- rank-r factorisation
struct LowrankBufs {
  float *P, *Po, *Z, *Zo, *G, *Linv, *B, *work, *evals, *V;
  int* status;  // [0] cholesky retries (max), [1] jacobi sweeps
  const uint64_t* seed_dev = nullptr;   // optional per-layer seed in device memory, added to the host seed
  SplitWs sw;   // split-K scratch shared by every contraction of the layer (they all run on one stream)
};


static LowrankBufs plan_lowrank(Arena& a, int64_t m, int64_t n, int64_t q, float* Zo_persistent, int* status) {
  LowrankBufs b;
  b.P = a.take<float>(n * q);
  b.Po = a.take<float>(n * q);
  b.Z = a.take<float>(m * q);
  b.Zo = Zo_persistent != nullptr ? Zo_persistent : a.take<float>(m * q);
  b.G = a.take<float>(q * q);
  b.Linv = a.take<float>(q * q);
  b.B = a.take<float>(q * n);
  b.work = a.take<float>(q * q + q + 8);
  b.evals = a.take<float>(q);
  b.V = a.take<float>(q * q);
  b.sw.buf = a.take<float>(kSplitWsBytes / sizeof(float)); b.sw.bytes = kSplitWsBytes;
  b.status = status;
  return b;
}

// X (N x q) -> Xo (N x q) with orthonormal columns: CholeskyQR, G = X^T X = Lc Lc^T, Xo = X Lc^-T
static int orthonormalize(const float* X, int64_t N, int64_t q, float* Xo, const LowrankBufs& b, cudaStream_t st) {
  CB_TRY(sgemm(q, q, N, 1.f, X, 1, q, X, q, 1, b.G, q, 1, false, nullptr, st, &b.sw));
  CB_TRY(cholesky_inverse(b.G, (int)q, b.Linv, b.status, st));
  CB_TRY(sgemm(N, q, q, 1.f, X, q, 1, b.Linv, 1, q, Xo, q, 1, false, nullptr, st, &b.sw));
  return CB_OK;
}

// Y: m x n (already column-weighted when aware).  On exit L (m x r), R (r x n).
static int lowrank_core(const float* Y, int64_t m, int64_t n, int64_t r, int64_t q, int niter, uint64_t seed,
                        int aware, const float* inv_sqrt_h, bool warm_valid, float* L, float* R,
                        const LowrankBufs& b, cudaStream_t st) {
  if (!warm_valid) {
    CB_TRY(fill_randn(b.P, n * q, seed, b.seed_dev, st));
    CB_TRY(sgemm(m, q, n, 1.f, Y, n, 1, b.P, q, 1, b.Z, q, 1, false, nullptr, st, &b.sw));
    CB_TRY(orthonormalize(b.Z, m, q, b.Zo, b, st));
  }
  for (int it = 0; it < niter; ++it) {
    CB_TRY(sgemm(n, q, m, 1.f, Y, 1, n, b.Zo, q, 1, b.P, q, 1, false, nullptr, st, &b.sw));  // P = Y^T Zo
    CB_TRY(orthonormalize(b.P, n, q, b.Po, b, st));
    CB_TRY(sgemm(m, q, n, 1.f, Y, n, 1, b.Po, q, 1, b.Z, q, 1, false, nullptr, st, &b.sw));  // Z = Y Po
    CB_TRY(orthonormalize(b.Z, m, q, b.Zo, b, st));
  }
  // second pass (CholeskyQR2) so the Rayleigh-Ritz basis is orthonormal to fp32 accuracy
  CB_TRY(orthonormalize(b.Zo, m, q, b.Z, b, st));
  CB_CUDA(cudaMemcpyAsync(b.Zo, b.Z, sizeof(float) * m * q, cudaMemcpyDeviceToDevice, st));
  // B = Zo^T Y (q x n); G = B B^T; eigen-decomposition through the Cholesky factor
  CB_TRY(sgemm(q, n, m, 1.f, b.Zo, 1, q, Y, n, 1, b.B, n, 1, false, nullptr, st, &b.sw));
  CB_TRY(sgemm(q, q, n, 1.f, b.B, n, 1, b.B, 1, n, b.G, q, 1, false, nullptr, st, &b.sw));
  CB_TRY(cholesky_inverse(b.G, (int)q, nullptr, b.status, st));
  CB_TRY(jacobi_eigh_from_chol(b.G, (int)q, b.evals, b.V, b.work, b.status != nullptr ? b.status + 1 : nullptr, st));
  // L = Zo V_r^T, R = V_r B (column-scaled by 1/sqrt(h) when aware: S V^T H^-1/2, alg.py:220-225)
  CB_TRY(sgemm(m, r, q, 1.f, b.Zo, q, 1, b.V, 1, q, L, r, 1, false, nullptr, st, &b.sw));
  CB_TRY(sgemm(r, n, q, 1.f, b.V, q, 1, b.B, n, 1, R, n, 1, false, aware ? inv_sqrt_h : nullptr, st, &b.sw));
  if (!aware) {
    // L = U sqrt(S), R = sqrt(S) V^T (alg.py:233-234); evals are squared singular values
    CB_TRY(scale_cols(L, m, r, b.evals, 2, L, st));
    CB_TRY(scale_rows(R, r, n, b.evals, 3, R, st));
  }
  // leave the basis rotated into its Ritz vectors (sorted): the next warm-started solve then
  // sees a nearly diagonal projected matrix and the Jacobi step converges in 2-3 sweeps
  CB_TRY(sgemm(m, q, q, 1.f, b.Zo, q, 1, b.V, 1, q, b.Z, q, 1, false, nullptr, st, &b.sw));
  CB_CUDA(cudaMemcpyAsync(b.Zo, b.Z, sizeof(float) * m * q, cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

// ---------------------------------------------------------------- tensor-core variant
// Same algorithm with every large contraction on the tcgen05 kernel (gemm_tc.cu).  All
// operands are bf16 and K-major, so each intermediate is kept in the orientation(s) its
// consumers contract over: "t" buffers are (q x long), plain ones (long x q).  Gram
// matrices, Cholesky factors and the Rayleigh-Ritz eigenproblem stay in fp32.
typedef __nv_bfloat16 bf16;
struct LowrankTcBufs {
  bf16 *Yb, *Ytb;                 // m x n, n x m
  bf16 *Ptb, *Pb, *Potb;          // q x n, n x q, q x n
  bf16 *Ztb, *Zb, *Zotb, *Zob;    // q x m, m x q, q x m (persistent for warm starts), m x q
  bf16 *Linvb, *Bb, *Btb, *Vb;    // q x q, q x n, n x q, q x q
  float *G, *Linv, *work, *evals, *V;
  int* status;                    // [0] cholesky retries, [1] jacobi sweeps, [2] gemm watchdog
  const uint64_t* seed_dev = nullptr;
  SplitWs sw;                     // the layer's split-K scratch (same one as LowrankBufs::sw)
  int* tile_counter = nullptr;    // dynamic tile scheduler of the batched contraction (2 zeroed ints)
};

static bool lowrank_tc_usable(int64_t m, int64_t n, int64_t r, int64_t q) {
  return m % 8 == 0 && n % 8 == 0 && q % 8 == 0 && r % 8 == 0 && m >= 256 && n >= 256 && q >= 16;
}

static LowrankTcBufs plan_lowrank_tc(Arena& a, int64_t m, int64_t n, int64_t q, int* status, const SplitWs& sw) {
  LowrankTcBufs b;
  b.Yb = a.take<bf16>(m * n); b.Ytb = a.take<bf16>(n * m);
  b.Ptb = a.take<bf16>(q * n); b.Pb = a.take<bf16>(n * q); b.Potb = a.take<bf16>(q * n);
  b.Ztb = a.take<bf16>(q * m); b.Zb = a.take<bf16>(m * q); b.Zotb = a.take<bf16>(q * m); b.Zob = a.take<bf16>(m * q);
  b.Linvb = a.take<bf16>(q * q); b.Bb = a.take<bf16>(q * n); b.Btb = a.take<bf16>(n * q); b.Vb = a.take<bf16>(q * q);
  b.G = a.take<float>(q * q); b.Linv = a.take<float>(q * q); b.work = a.take<float>(q * q + q + 8);
  b.evals = a.take<float>(q); b.V = a.take<float>(q * q);
  b.tile_counter = a.take<int>(4);
  b.sw = sw;
  b.status = status;
  return b;
}

__global__ void __launch_bounds__(256) randn_bf16_kernel(bf16* __restrict__ p_, int64_t count, uint64_t seed,
                                                         const uint64_t* __restrict__ seed_dev, int64_t bstride) {
  bf16* __restrict__ p = boff(p_, bstride * blockIdx.y);           // one layer of a batch per blockIdx.y
  if (seed_dev != nullptr) seed += boff(seed_dev, bstride * blockIdx.y)[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    p[i] = __float2bfloat16_rn(gaussian_from(seed, (uint64_t)i));
}

// Xt (q x N) and X (N x q) hold the same raw sketch; writes the orthonormalised sketch as
// Xot (q x N) and/or Xo (N x q).
static int orthonormalize_tc(const bf16* Xt, const bf16* X, int64_t N, int64_t q, bf16* Xot, bf16* Xo,
                             const LowrankTcBufs& b, cudaStream_t st) {
  int* wd = b.status != nullptr ? b.status + 2 : nullptr;
  CB_TRY(gemm_tc(q, q, N, 1.f, Xt, N, Xt, N, b.G, q, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, wd, nullptr, st, &b.sw));
  CB_TRY(cholesky_inverse(b.G, (int)q, b.Linv, b.status, st, b.Linvb));
  // Xot[q, N] = Linv[q, K=q] * X[N, K=q]^T
  CB_TRY(gemm_tc(q, N, q, 1.f, b.Linvb, q, X, q, nullptr, 0, Xot, N, Xo, q, nullptr, nullptr, 1, wd, nullptr, st));
  return CB_OK;
}

static int lowrank_core_tc(int64_t m, int64_t n, int64_t r, int64_t q, int niter, uint64_t seed, int aware,
                           const float* inv_sqrt_h, bool warm_valid, float* L, float* R, bf16* Lb, bf16* Rtb,
                           const LowrankTcBufs& b, cudaStream_t st) {
  int* wd = b.status != nullptr ? b.status + 2 : nullptr;
  if (!warm_valid) {
    randn_bf16_kernel<<<grid_for(q * n, 256 * 4, 4), 256, 0, st>>>(b.Ptb, q * n, seed, b.seed_dev, 0);
    CB_CHECK_LAUNCH();
    // Zt[q, m] = Pt[q, K=n] * Y[m, K=n]^T
    CB_TRY(gemm_tc(q, m, n, 1.f, b.Ptb, n, b.Yb, n, nullptr, 0, b.Ztb, m, b.Zb, q, nullptr, nullptr, 1, wd, nullptr, st));
    CB_TRY(orthonormalize_tc(b.Ztb, b.Zb, m, q, b.Zotb, b.Zob, b, st));
  }
  for (int it = 0; it < niter; ++it) {
    // Pt[q, n] = Zot[q, K=m] * Yt[n, K=m]^T
    CB_TRY(gemm_tc(q, n, m, 1.f, b.Zotb, m, b.Ytb, m, nullptr, 0, b.Ptb, n, b.Pb, q, nullptr, nullptr, 1, wd, nullptr, st));
    CB_TRY(orthonormalize_tc(b.Ptb, b.Pb, n, q, b.Potb, nullptr, b, st));
    CB_TRY(gemm_tc(q, m, n, 1.f, b.Potb, n, b.Yb, n, nullptr, 0, b.Ztb, m, b.Zb, q, nullptr, nullptr, 1, wd, nullptr, st));
    CB_TRY(orthonormalize_tc(b.Ztb, b.Zb, m, q, b.Zotb, b.Zob, b, st));
  }
  // CholeskyQR2 on the (bf16-rounded) basis itself
  CB_CUDA(cudaMemcpyAsync(b.Ztb, b.Zotb, sizeof(bf16) * q * m, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaMemcpyAsync(b.Zb, b.Zob, sizeof(bf16) * q * m, cudaMemcpyDeviceToDevice, st));
  CB_TRY(orthonormalize_tc(b.Ztb, b.Zb, m, q, b.Zotb, b.Zob, b, st));
  // B[q, n] = Zot[q, K=m] * Yt[n, K=m]^T  (bf16 in both orientations); G = B B^T in fp32
  CB_TRY(gemm_tc(q, n, m, 1.f, b.Zotb, m, b.Ytb, m, nullptr, 0, b.Bb, n, b.Btb, q, nullptr, nullptr, 1, wd, nullptr, st));
  CB_TRY(gemm_tc(q, q, n, 1.f, b.Bb, n, b.Bb, n, b.G, q, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, wd, nullptr, st, &b.sw));
  CB_TRY(cholesky_inverse(b.G, (int)q, nullptr, b.status, st));
  CB_TRY(jacobi_eigh_from_chol(b.G, (int)q, b.evals, b.V, b.work, b.status != nullptr ? b.status + 1 : nullptr, st));
  CB_TRY(to_bf16(b.V, q, q, q, b.Vb, q, nullptr, 0, nullptr, st));
  // L[m, r] = Zo[m, K=q] * V_r[r, K=q]^T ;  R[r, n] = V_r[r, K=q] * Bt[n, K=q]^T (col-scaled by 1/sqrt(h))
  CB_TRY(gemm_tc(m, r, q, 1.f, b.Zob, q, b.Vb, q, L, r, aware ? Lb : nullptr, r, nullptr, 0, nullptr, nullptr, 1, wd, nullptr, st));
  CB_TRY(gemm_tc(r, n, q, 1.f, b.Vb, q, b.Btb, q, R, n, nullptr, 0, aware ? Rtb : nullptr, r, aware ? inv_sqrt_h : nullptr,
                 nullptr, 1, wd, nullptr, st));
  if (!aware) {
    CB_TRY(scale_cols(L, m, r, b.evals, 2, L, st));
    CB_TRY(scale_rows(R, r, n, b.evals, 3, R, st));
  }
  // rotate the stored basis into its Ritz vectors for the next warm start:
  // Zrot[m, q] = Zo[m, K=q] * V[q, K=q]^T, written in both orientations
  CB_TRY(gemm_tc(m, q, q, 1.f, b.Zob, q, b.Vb, q, nullptr, 0, b.Zb, q, b.Ztb, m, nullptr, nullptr, 1, wd, nullptr, st));
  CB_CUDA(cudaMemcpyAsync(b.Zob, b.Zb, sizeof(bf16) * m * q, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaMemcpyAsync(b.Zotb, b.Ztb, sizeof(bf16) * m * q, cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

__global__ void sqrt_vec_kernel(const float* in, int n, float* out) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sqrtf(fmaxf(in[i], 0.f));
}

static int64_t default_sketch_width(const cb_caldera_params* p, int64_t m, int64_t n) {
  const int64_t mn = m < n ? m : n;
  int64_t q;
  if (p->sketch_width > 0) q = p->sketch_width;
  else if (p->rand_svd) q = 2 * (int64_t)p->rank;                               // alg.py:213
  else q = (2 * (int64_t)p->rank > (int64_t)p->rank + 32) ? 2 * (int64_t)p->rank : (int64_t)p->rank + 32;
  // the shared-memory Rayleigh-Ritz solver (smalldense.cu) holds q <= 224 columns; prefer it
  // whenever that still leaves >= 32 columns of oversampling
  if (p->sketch_width <= 0 && q > 224 && (int64_t)p->rank + 32 <= 224) q = 224;
  if (q > mn) q = mn;
  if (q > 512) q = 512;
  if (q < p->rank) q = p->rank;
  return q;
}

// ---------------------------------------------------------------- layer plan
struct LayerPlan {
  double* dsc;     // [0] sumsq [1] den [2] num [3] num_inner
  int* flags;      // [0] take_outer [1] take_inner [2] cholesky retries [3] jacobi sweeps
  float* scalars;  // 8
  float *h_eff, *sqrt_h, *inv_sqrt_h, *w_inner, *amax;
  float* Ws;
  void* codes_cur;
  float* qscale_cur;
  float *LRbuf, *Y, *RES, *Lcur, *Rcur, *Zwarm;
  // LPLR
  float *Rw, *Gs, *Linv, *Ginv, *Bl, *Br, *Ltmp, *Rtmp, *Lb, *Rb;
  void *Lcodes_cur, *Rcodes_cur, *Lcodes_in, *Rcodes_in, *Lcodes_out;
  float *Lscale_cur, *Rscale_cur, *Lscale_in, *Rscale_in;
  LowrankBufs lr;
  LowrankTcBufs tc;
  bf16 *Lb16, *Rtb16;   // bf16 hi/lo splits of L (m x 3r) and R^T (n x 3r) for the tensor-core L R product
  bf16 *Rsb16, *Ltb16;  // LPLR operands: R (.) sqrt(h) (r x n) and L^T (r x m)
  // dense (non-diagonal) Hessian
  float *Hs, *Ebuf, *Tbuf, *HP, *HRt, *eigwork;
  int64_t q;
  bool quant_factors;
  bool use_tc;
  bool dense;
};

static int plan_layer(Arena& a, const cb_caldera_params* p, int64_t m, int64_t n, bool scale_w, int h_kind,
                      LayerPlan& L) {
  const int64_t r = p->rank;
  L.dense = (h_kind == CB_H_DENSE);
  L.quant_factors = p->compute_lr && (p->l_bits < 16 || p->r_bits < 16);
  L.q = p->compute_lr ? default_sketch_width(p, m, n) : 0;
  L.dsc = a.take<double>(4);
  L.flags = a.take<int>(8);
  L.scalars = a.take<float>(8);
  L.h_eff = a.take<float>(n);
  L.sqrt_h = a.take<float>(n);
  L.inv_sqrt_h = a.take<float>(n);
  L.w_inner = a.take<float>(n);
  L.amax = a.take<float>(4);
  L.Ws = scale_w ? a.take<float>(m * n) : nullptr;
  L.codes_cur = p->compute_q ? (void*)a.take<uint8_t>(m * n * code_bytes(p->q_bits)) : nullptr;
  L.qscale_cur = a.take<float>(4);
  if (p->compute_lr) {
    L.LRbuf = a.take<float>(m * n);
    L.Y = a.take<float>(m * n);
    L.RES = (L.quant_factors && p->aware && !L.dense) ? a.take<float>(m * n) : nullptr;
    L.Lcur = a.take<float>(m * r);
    L.Rcur = a.take<float>(r * n);
    L.Zwarm = a.take<float>(m * L.q);
    L.lr = plan_lowrank(a, m, n, L.q, L.Zwarm, nullptr);
    L.use_tc = p->use_tensor_cores != 0 && !L.dense && lowrank_tc_usable(m, n, r, L.q);
    if (L.use_tc) {
      L.tc = plan_lowrank_tc(a, m, n, L.q, nullptr, L.lr.sw);
      L.Lb16 = a.take<bf16>(3 * m * r);
      L.Rtb16 = a.take<bf16>(3 * n * r);
      if (L.quant_factors) {
        L.Rsb16 = a.take<bf16>(r * n);
        L.Ltb16 = a.take<bf16>(r * m);
      }
    }
    if (L.quant_factors) {
      L.Rw = a.take<float>(r * n);
      L.Gs = a.take<float>(r * r);
      L.Linv = a.take<float>(r * r);
      L.Ginv = a.take<float>(r * r);
      L.Bl = a.take<float>(m * r);
      L.Br = a.take<float>(r * n);
      L.Ltmp = a.take<float>(m * r);
      L.Rtmp = a.take<float>(r * n);
      L.Lb = a.take<float>(m * r);
      L.Rb = a.take<float>(r * n);
      L.Lcodes_cur = a.take<uint8_t>(m * r * code_bytes(p->l_bits));
      L.Rcodes_cur = a.take<uint8_t>(r * n * code_bytes(p->r_bits));
      L.Lcodes_in = a.take<uint8_t>(m * r * code_bytes(p->l_bits));
      L.Rcodes_in = a.take<uint8_t>(r * n * code_bytes(p->r_bits));
      L.Lcodes_out = a.take<uint8_t>(m * r * code_bytes(p->l_bits));
      L.Lscale_cur = a.take<float>(4);
      L.Rscale_cur = a.take<float>(4);
      L.Lscale_in = a.take<float>(4);
      L.Rscale_in = a.take<float>(4);
    }
  }
  if (L.dense) {
    L.Hs = a.take<float>(n * n);
    L.eigwork = p->aware ? a.take<float>(min_eig_shift_workspace_floats(n)) : nullptr;
    L.Ebuf = a.take<float>(m * n);
    L.Tbuf = a.take<float>(m * n);
    if (p->compute_lr) {
      L.HP = a.take<float>(n * L.q);
      L.HRt = a.take<float>(n * r);
    }
  }
  return a.ok() ? CB_OK : CB_ERR_WORKSPACE;
}

static int validate_params(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind) {
  if (p == nullptr || m <= 0 || n <= 0) return CB_ERR_ARG;
  if (p->n_order < 0 || p->n_order > 8 || p->iters < 0) return CB_ERR_ARG;
  if (p->exec_mode != CB_MODE_LATENCY && p->exec_mode != CB_MODE_THROUGHPUT) return CB_ERR_ARG;
  for (int i = 0; i < p->n_order; ++i)
    if (p->order[i] != 0 && p->order[i] != 1) return CB_ERR_ARG;
  if (p->compute_q && !bits_ok(p->q_bits)) return CB_ERR_BITS;
  if (p->compute_lr) {
    if (!bits_ok(p->l_bits) || !bits_ok(p->r_bits)) return CB_ERR_BITS;
    if (p->rank < 1 || p->rank > m || p->rank > n) return CB_ERR_ARG;
    if (p->rank > 512) return CB_ERR_UNSUPPORTED;
    // lplr_iters < 1 with quantised factors only fails in the reference when an LR update actually runs
    // (best_L_quant_out stays None, alg.py:190); update_order = ["Q"] or iters = 0 are fine there and here
    bool lr_scheduled = false;
    for (int i = 0; i < p->n_order; ++i) lr_scheduled = lr_scheduled || p->order[i] == 1;
    if ((p->l_bits < 16 || p->r_bits < 16) && p->lplr_iters < 1 && lr_scheduled && p->iters > 0) return CB_ERR_ARG;
  }
  if (h_kind != CB_H_IDENTITY && h_kind != CB_H_DIAG && h_kind != CB_H_DENSE) return CB_ERR_ARG;
  if (p->q_block != 0) return CB_ERR_UNSUPPORTED;
  return CB_OK;
}

// whole-tensor quantise + dequantise (quantize_matrix, alg.py:245-250)
static int quantize_whole(const float* x, int64_t rows, int64_t cols, int bits, void* codes, float* scale,
                          float* deq, cudaStream_t st) {
  return cb_quantize_f32(x, rows, cols, cols, 1, bits, 0, 1e-8f, codes, nullptr, scale, deq, (void*)st);
}

static int solve_spd_setup(float* G, int64_t r, float* Linv, float* Ginv, int* status, cudaStream_t st) {
  CB_TRY(cholesky_inverse(G, (int)r, Linv, status, st));
  // G^-1 = Linv^T Linv
  CB_TRY(sgemm(r, r, r, 1.f, Linv, 1, r, Linv, r, 1, Ginv, r, 1, false, nullptr, st));
  return CB_OK;
}

// ---------------------------------------------------------------- dense Hessian
// For a general symmetric H the weighted problem min ||(res - L R) H^(1/2)||_F needs no square
// root of H at all: the left singular vectors of res H^(1/2) are the eigenvectors of
// A = res H res^T, so the subspace iteration applies A as three products
//   P = res^T Z,   HP = H P,   Z' = res HP
// (P is re-orthonormalised in between, which does not change the span), Rayleigh-Ritz uses
// G = P^T H P with P = res^T Zo, and R = L^T res = V_r P^T (alg.py:210-225 with V V^T = I).
static int lowrank_core_dense(const float* res, const float* Hs, int64_t m, int64_t n, int64_t r, int64_t q,
                              int niter, uint64_t seed, bool warm_valid, float* L, float* R, float* HP,
                              const LowrankBufs& b, cudaStream_t st) {
  if (!warm_valid) {
    CB_TRY(fill_randn(b.P, n * q, seed, b.seed_dev, st));
    CB_TRY(sgemm(n, q, n, 1.f, Hs, n, 1, b.P, q, 1, HP, q, 1, false, nullptr, st, &b.sw));
    CB_TRY(sgemm(m, q, n, 1.f, res, n, 1, HP, q, 1, b.Z, q, 1, false, nullptr, st, &b.sw));
    CB_TRY(orthonormalize(b.Z, m, q, b.Zo, b, st));
  }
  for (int it = 0; it < niter; ++it) {
    CB_TRY(sgemm(n, q, m, 1.f, res, 1, n, b.Zo, q, 1, b.P, q, 1, false, nullptr, st, &b.sw));   // P = res^T Zo
    CB_TRY(orthonormalize(b.P, n, q, b.Po, b, st));
    CB_TRY(sgemm(n, q, n, 1.f, Hs, n, 1, b.Po, q, 1, HP, q, 1, false, nullptr, st, &b.sw));     // HP = H Po
    CB_TRY(sgemm(m, q, n, 1.f, res, n, 1, HP, q, 1, b.Z, q, 1, false, nullptr, st, &b.sw));     // Z = res HP
    CB_TRY(orthonormalize(b.Z, m, q, b.Zo, b, st));
  }
  CB_TRY(orthonormalize(b.Zo, m, q, b.Z, b, st));
  CB_CUDA(cudaMemcpyAsync(b.Zo, b.Z, sizeof(float) * m * q, cudaMemcpyDeviceToDevice, st));
  CB_TRY(sgemm(n, q, m, 1.f, res, 1, n, b.Zo, q, 1, b.P, q, 1, false, nullptr, st, &b.sw));     // P = res^T Zo
  CB_TRY(sgemm(n, q, n, 1.f, Hs, n, 1, b.P, q, 1, HP, q, 1, false, nullptr, st, &b.sw));
  CB_TRY(sgemm(q, q, n, 1.f, b.P, 1, q, HP, q, 1, b.G, q, 1, false, nullptr, st, &b.sw));       // G = P^T H P
  CB_TRY(cholesky_inverse(b.G, (int)q, nullptr, b.status, st));
  CB_TRY(jacobi_eigh_from_chol(b.G, (int)q, b.evals, b.V, b.work, b.status != nullptr ? b.status + 1 : nullptr, st));
  CB_TRY(sgemm(m, r, q, 1.f, b.Zo, q, 1, b.V, 1, q, L, r, 1, false, nullptr, st, &b.sw));       // L = Zo V_r^T
  CB_TRY(sgemm(r, n, q, 1.f, b.V, q, 1, b.P, 1, q, R, n, 1, false, nullptr, st, &b.sw));        // R = V_r P^T
  CB_TRY(sgemm(m, q, q, 1.f, b.Zo, q, 1, b.V, 1, q, b.Z, q, 1, false, nullptr, st, &b.sw));     // Ritz rotation (warm start)
  CB_CUDA(cudaMemcpyAsync(b.Zo, b.Z, sizeof(float) * m * q, cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

// out += tr(E H E^T) (squared == false) or ||E H||_F^2 (squared == true), E = A - Q - LR
static int dense_quadratic(const LayerPlan& P, const float* A, const void* codes, int bits, const float* qscale,
                           const float* LR, int64_t m, int64_t n, bool squared, double* out, cudaStream_t st) {
  const float* E = A;
  if (codes != nullptr || LR != nullptr) {
    CB_TRY(form_e(A, codes, bits, qscale, LR, m, n, P.Ebuf, st));
    E = P.Ebuf;
  }
  CB_TRY(sgemm(m, n, n, 1.f, E, n, 1, P.Hs, n, 1, P.Tbuf, n, 1, false, nullptr, st, &P.lr.sw));
  return dot_accum(P.Tbuf, squared ? P.Tbuf : E, m * n, out, st);
}

// LRbuf = L R (m x n fp32): tensor cores when the shape allows, SIMT otherwise.  The factors
// are split into bf16 hi + lo parts and contracted as one K = 3r GEMM
//   [Lh | Lh | Ll] [Rh | Rl | Rh]^T = Lh Rh + Lh Rl + Ll Rh
// so the product carries ~16 mantissa bits: the residual W - L R that the Q update quantises
// and the reported error then agree with the fp32 factors that are returned.
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ X_, int64_t rows, int64_t cols, int transpose, int second_is_lo,
              bf16* __restrict__ out_ /* rows_out x 3*inner */, int64_t bstride = 0) {
  const float* __restrict__ X = boff(X_, bstride * blockIdx.y);
  bf16* __restrict__ out = boff(out_, bstride * blockIdx.y);
  // transpose == 0: X is rows x cols, out row i = [hi(i,:) | (lo or hi)(i,:) | (hi or lo)(i,:)] with inner = cols
  // transpose == 1: X is rows x cols, out row j (of cols) built from column j, inner = rows
  const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t inner = transpose ? rows : cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / cols, j = e - i * cols;
    const float v = X[e];
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int64_t orow = transpose ? j : i, ok = transpose ? i : j;
    bf16* o = out + orow * 3 * inner + ok;
    o[0] = hi;
    o[inner] = second_is_lo ? lo : hi;
    o[2 * inner] = second_is_lo ? hi : lo;
  }
}

// transpose == 1 form of split3_kernel through a 32 x 32 shared-memory tile: reads run along a row of X, writes along
// the inner index of an output row (blockIdx.z = layer of a batch).
__global__ void __launch_bounds__(256)
split3_t_kernel(const float* __restrict__ X_, int rows, int cols, int second_is_lo, bf16* __restrict__ out_, int64_t bstride) {
  __shared__ float tile[32][33];
  const float* __restrict__ X = boff(X_, bstride * blockIdx.z);
  bf16* __restrict__ out = boff(out_, bstride * blockIdx.z);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, j = j0 + tx;
    tile[ty + 8 * k][tx] = (i < rows && j < cols) ? X[(int64_t)i * cols + j] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = j0 + ty + 8 * k, i = i0 + tx;
    if (i < rows && j < cols) {
      const float v = tile[tx][ty + 8 * k];
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      bf16* o = out + (int64_t)j * 3 * rows + i;
      o[0] = hi;
      o[rows] = second_is_lo ? lo : hi;
      o[2 * (int64_t)rows] = second_is_lo ? hi : lo;
    }
  }
}

static int lr_product_raw(bool use_tc, const float* Lcur, const float* Rcur, int64_t m, int64_t n, int64_t r,
                          float* LRbuf, bf16* Lb16, bf16* Rtb16, int* watchdog, cudaStream_t st) {
  if (use_tc) {
    // A' = [Lh | Lh | Ll] (m x 3r),  B' = [Rh^T | Rl^T | Rh^T] (n x 3r)
    split3_kernel<<<grid_for(m * r, 256 * 4, 4), 256, 0, st>>>(Lcur, m, r, 0, 0, Lb16);
    CB_CHECK_LAUNCH();
    split3_t_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((r + 31) / 32), 1), 256, 0, st>>>(Rcur, (int)r, (int)n, 1, Rtb16, 0);
    CB_CHECK_LAUNCH();
    return gemm_tc(m, n, 3 * r, 1.f, Lb16, 3 * r, Rtb16, 3 * r, LRbuf, n, nullptr, 0, nullptr, 0, nullptr, nullptr, 1,
                   watchdog, nullptr, st);
  }
  return sgemm(m, n, r, 1.f, Lcur, r, 1, Rcur, n, 1, LRbuf, n, 1, false, nullptr, st);
}

static int lr_product(const LayerPlan& P, int64_t m, int64_t n, int64_t r, cudaStream_t st) {
  return lr_product_raw(P.use_tc, P.Lcur, P.Rcur, m, n, r, P.LRbuf, P.Lb16, P.Rtb16, P.flags + 4, st);
}

// One LPLR iteration (alg.py:161-182) from the current R: weighted least squares for L, quantise L^T, least
// squares for R, quantise R, inner error sum w_j (res - L R)_ij^2 accumulated into P.dsc[3].  On exit
// P.Ltmp / P.Rtmp hold the unquantised solutions, P.Lcodes_cur / P.Rcodes_cur + scales the codes (L's in m x r
// order; the reference's L_idxs order is the transpose), P.Lcur / P.Rcur the dequantised factors.
static int lplr_step(const cb_caldera_params* p, const LayerPlan& P, int64_t m, int64_t n, cudaStream_t st) {
  const int64_t r = p->rank;
  const float* res = (p->aware && !P.dense) ? P.RES : P.Y;  // unweighted residual W - Q
  // ---- L update: weighted normal equations (alg.py:163 / :167)
  if (P.dense && p->aware) {
    CB_TRY(sgemm(n, r, n, 1.f, P.Hs, n, 1, P.Rcur, 1, n, P.HRt, r, 1, false, nullptr, st, &P.lr.sw));   // H R^T
    CB_TRY(sgemm(r, r, n, 1.f, P.Rcur, n, 1, P.HRt, r, 1, P.Gs, r, 1, false, nullptr, st, &P.lr.sw));   // R H R^T
    CB_TRY(sgemm(m, r, n, 1.f, res, n, 1, P.HRt, r, 1, P.Bl, r, 1, false, nullptr, st, &P.lr.sw));      // res H R^T
  } else {
    const float* Rw = P.Rcur;
    if (p->aware) { CB_TRY(scale_cols(P.Rcur, r, n, P.h_eff, 0, P.Rw, st)); Rw = P.Rw; }
    CB_TRY(sgemm(r, r, n, 1.f, Rw, n, 1, P.Rcur, 1, n, P.Gs, r, 1, false, nullptr, st, &P.lr.sw));     // R diag(h) R^T
    CB_TRY(sgemm(m, r, n, 1.f, res, n, 1, Rw, 1, n, P.Bl, r, 1, false, nullptr, st, &P.lr.sw));        // res diag(h) R^T
  }
  CB_TRY(solve_spd_setup(P.Gs, r, P.Linv, P.Ginv, P.flags + 2, st));
  CB_TRY(sgemm(m, r, r, 1.f, P.Bl, r, 1, P.Ginv, r, 1, P.Ltmp, r, 1, false, nullptr, st, &P.lr.sw));
  CB_TRY(quantize_whole(P.Ltmp, m, r, p->l_bits, P.Lcodes_cur, P.Lscale_cur, P.Lcur, st));  // alg.py:171-172
  // ---- R update (alg.py:175)
  CB_TRY(sgemm(r, r, m, 1.f, P.Lcur, 1, r, P.Lcur, r, 1, P.Gs, r, 1, false, nullptr, st, &P.lr.sw));  // L^T L
  CB_TRY(sgemm(r, n, m, 1.f, P.Lcur, 1, r, res, n, 1, P.Br, n, 1, false, nullptr, st, &P.lr.sw));     // L^T res
  CB_TRY(solve_spd_setup(P.Gs, r, P.Linv, P.Ginv, P.flags + 2, st));
  CB_TRY(sgemm(r, n, r, 1.f, P.Ginv, r, 1, P.Br, n, 1, P.Rtmp, n, 1, false, nullptr, st, &P.lr.sw));
  CB_TRY(quantize_whole(P.Rtmp, r, n, p->r_bits, P.Rcodes_cur, P.Rscale_cur, P.Rcur, st));  // alg.py:179-180
  // ---- inner error ||(res - L R) H_sqrt||_F (alg.py:182)
  CB_TRY(sgemm(m, n, r, 1.f, P.Lcur, r, 1, P.Rcur, n, 1, P.LRbuf, n, 1, false, nullptr, st, &P.lr.sw));
  if (P.dense) CB_TRY(dense_quadratic(P, res, nullptr, 8, nullptr, P.LRbuf, m, n, !p->aware, P.dsc + 3, st));
  else CB_TRY(err_accum(res, nullptr, 8, nullptr, P.LRbuf, P.w_inner, m, n, P.dsc + 3, st));
  return CB_OK;
}

// Tensor-core variant (diagonal / identity Hessian): the three m x n x r contractions of an iteration run on
// gemm_tc against the bf16 operands Yb = res (.) sqrt(h) and Ytb = Yb^T that the rank-r step already built:
//   res diag(h) R^T = Yb (R (.) sqrt(h))^T,        L^T res = (L^T Ytb^T) (.) 1/sqrt(h)
// The r x r Gram matrices are accumulated in fp32 (split-K slices summed in slice order) and factorised in fp32.
static int lplr_step_tc(const cb_caldera_params* p, const LayerPlan& P, int64_t m, int64_t n, cudaStream_t st) {
  const int64_t r = p->rank;
  const float* res = p->aware ? P.RES : P.Y;
  int* wd = P.flags + 4;
  // ---- L update (alg.py:163 / :167)
  CB_TRY(to_bf16(P.Rcur, r, n, n, P.Rsb16, n, nullptr, 0, p->aware ? P.sqrt_h : nullptr, st));
  CB_TRY(gemm_tc(r, r, n, 1.f, P.Rsb16, n, P.Rsb16, n, P.Gs, r, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, wd, nullptr, st, &P.tc.sw));
  CB_TRY(gemm_tc(m, r, n, 1.f, P.tc.Yb, n, P.Rsb16, n, P.Bl, r, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, wd, nullptr, st, &P.tc.sw));
  CB_TRY(solve_spd_setup(P.Gs, r, P.Linv, P.Ginv, P.flags + 2, st));
  CB_TRY(sgemm(m, r, r, 1.f, P.Bl, r, 1, P.Ginv, r, 1, P.Ltmp, r, 1, false, nullptr, st, &P.lr.sw));
  CB_TRY(quantize_whole(P.Ltmp, m, r, p->l_bits, P.Lcodes_cur, P.Lscale_cur, P.Lcur, st));
  // ---- R update (alg.py:175)
  CB_TRY(to_bf16(P.Lcur, m, r, r, nullptr, 0, P.Ltb16, m, nullptr, st));
  CB_TRY(gemm_tc(r, r, m, 1.f, P.Ltb16, m, P.Ltb16, m, P.Gs, r, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, wd, nullptr, st, &P.tc.sw));
  CB_TRY(gemm_tc(r, n, m, 1.f, P.Ltb16, m, P.tc.Ytb, m, P.Br, n, nullptr, 0, nullptr, 0, p->aware ? P.inv_sqrt_h : nullptr,
                 nullptr, 0, wd, nullptr, st, &P.tc.sw));
  CB_TRY(solve_spd_setup(P.Gs, r, P.Linv, P.Ginv, P.flags + 2, st));
  CB_TRY(sgemm(r, n, r, 1.f, P.Ginv, r, 1, P.Br, n, 1, P.Rtmp, n, 1, false, nullptr, st, &P.lr.sw));
  CB_TRY(quantize_whole(P.Rtmp, r, n, p->r_bits, P.Rcodes_cur, P.Rscale_cur, P.Rcur, st));
  // ---- inner error (alg.py:182)
  CB_TRY(lr_product(P, m, n, r, st));
  CB_TRY(err_accum(res, nullptr, 8, nullptr, P.LRbuf, P.w_inner, m, n, P.dsc + 3, st));
  return CB_OK;
}

// The LPLR loop (alg.py:160-195): lplr_iters steps, best-so-far kept by predicated device copies.
static int lplr_refine(const cb_caldera_params* p, const LayerPlan& P, int64_t m, int64_t n, bool use_tc, cudaStream_t st) {
  const int64_t r = p->rank;
  for (int k = 0; k < p->lplr_iters; ++k) {
    CB_TRY(use_tc ? lplr_step_tc(p, P, m, n, st) : lplr_step(p, P, m, n, st));
    CB_TRY(select_inner(P.dsc + 3, P.scalars, P.flags, k == 0, k == p->lplr_iters - 1, st));   // alg.py:184-188
    CopySegments best;
    best.add(P.Lb, P.Lcur, sizeof(float) * m * r);
    best.add(P.Rb, P.Rcur, sizeof(float) * r * n);
    best.add(P.Lcodes_in, P.Lcodes_cur, (size_t)m * r * code_bytes(p->l_bits));
    best.add(P.Rcodes_in, P.Rcodes_cur, (size_t)r * n * code_bytes(p->r_bits));
    best.add(P.Lscale_in, P.Lscale_cur, sizeof(float));
    best.add(P.Rscale_in, P.Rscale_cur, sizeof(float));
    CB_TRY(copy_if_multi(P.flags + 1, best, st));
  }
  CB_CUDA(cudaMemcpyAsync(P.Lcur, P.Lb, sizeof(float) * m * r, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaMemcpyAsync(P.Rcur, P.Rb, sizeof(float) * r * n, cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

// ================================================================ batched layers
// B same-shape layers advancing in lock step (Bt: layer b's buffers sit b * stride bytes after layer 0's).  Every
// contraction of the rank-r step, the small factorisations and the bookkeeping kernels are ONE launch for the whole
// batch -- the batched CTA-pair contraction of gemm_tc2.cu (hundreds of tiles per launch instead of 16-32), one CTA
// per layer for Cholesky / Jacobi -- and only the HBM-bound full-matrix passes, which fill the machine on their own,
// are launched layer by layer.  Orientation: gemm_tc2 owns 256-row tiles of its M, so the long dimension of every
// skinny contraction goes to M and the sketch width q (224) becomes the instruction's N.
static int g2(const Bt& bt, int64_t M, int64_t N, int64_t K, const bf16* A, int64_t lda, const bf16* B, int64_t ldb,
              float* C, int64_t ldc, bf16* Cb, int64_t ldcb, bf16* Ct, int64_t ldct, const float* colscale, int* wd,
              cudaStream_t st, int* tile_counter = nullptr, bf16* Cs = nullptr, int split_mode = 0) {
  Gemm2Batch g;
  g.tile_counter = tile_counter;
  g.Cs = Cs; g.sCs = bt.stride; g.split_mode = split_mode;
  g.batch = bt.n; g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.sA = bt.stride; g.B = B; g.ldb = ldb; g.sB = bt.stride;
  g.C = C; g.ldc = ldc; g.sC = bt.stride; g.Cb = Cb; g.ldcb = ldcb; g.sCb = bt.stride; g.Ct = Ct; g.ldct = ldct; g.sCt = bt.stride;
  g.colscale = colscale; g.sCol = bt.stride; g.error_flag = wd;
  return gemm_tc2(g, st);
}

// Xt (q x N) and X (N x q) hold the same raw sketch; writes its orthonormalised form as Xot (q x N) and/or Xo (N x q)
static int orthonormalize_b(const Bt& bt, const bf16* Xt, const bf16* X, int64_t N, int64_t q, bf16* Xot, bf16* Xo,
                            const LowrankTcBufs& b, cudaStream_t st) {
  int* wd = b.status != nullptr ? b.status + 2 : nullptr;
  CB_TRY(g2(bt, q, q, N, Xt, N, Xt, N, b.G, q, nullptr, 0, nullptr, 0, nullptr, wd, st, b.tile_counter));              // G = X^T X
  CB_TRY(cholesky_inverse(b.G, (int)q, b.Linv, b.status, st, b.Linvb, bt));
  // Xot[q, N] = Linv[q, K = q] * X[N, K = q]^T with the sketch width on M: the orientation every consumer inside the
  // power loop needs (Xot) is then the row-major output (whole 64-byte pieces per thread); the transposed one (Xo,
  // 2-byte stores) is only written when a caller asks for it.  K is tiny here, so the epilogue is the kernel.
  return g2(bt, q, N, q, b.Linvb, q, X, q, nullptr, 0, Xot, N, Xo, q, nullptr, wd, st, b.tile_counter);
}

static int copy_b(const Bt& bt, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  CopySegments c;
  c.add(dst, src, bytes);
  return copy_if_multi(nullptr, c, st, bt);
}

// Lsplit / Rtsplit (optional, aware only): the split-bf16 operands of the L R contraction, written by the epilogues of
// the two contractions that produce L and R instead of by a conversion pass of their own
static int lowrank_core_b(const Bt& bt, int64_t m, int64_t n, int64_t r, int64_t q, int niter, uint64_t seed, int aware,
                          const float* inv_sqrt_h, bool warm_valid, float* L, float* R, const LowrankTcBufs& b, cudaStream_t st,
                          bf16* Lsplit = nullptr, bf16* Rtsplit = nullptr) {
  int* wd = b.status != nullptr ? b.status + 2 : nullptr;
  if (!warm_valid) {
    dim3 grid((unsigned)grid_for(q * n, 256 * 4, 1), (unsigned)bt.n);
    randn_bf16_kernel<<<grid, 256, 0, st>>>(b.Ptb, q * n, seed, b.seed_dev, bt.stride);
    CB_CHECK_LAUNCH();
    CB_TRY(g2(bt, m, q, n, b.Yb, n, b.Ptb, n, nullptr, 0, b.Zb, q, b.Ztb, m, nullptr, wd, st, b.tile_counter));        // Z = Y P
    CB_TRY(orthonormalize_b(bt, b.Ztb, b.Zb, m, q, b.Zotb, niter == 0 ? b.Zob : nullptr, b, st));
  }
  for (int it = 0; it < niter; ++it) {
    CB_TRY(g2(bt, n, q, m, b.Ytb, m, b.Zotb, m, nullptr, 0, b.Pb, q, b.Ptb, n, nullptr, wd, st, b.tile_counter));      // P = Y^T Zo
    CB_TRY(orthonormalize_b(bt, b.Ptb, b.Pb, n, q, b.Potb, nullptr, b, st));
    CB_TRY(g2(bt, m, q, n, b.Yb, n, b.Potb, n, nullptr, 0, b.Zb, q, b.Ztb, m, nullptr, wd, st, b.tile_counter));       // Z = Y Po
    // (the m x q orientation of the basis is only consumed after the loop)
    CB_TRY(orthonormalize_b(bt, b.Ztb, b.Zb, m, q, b.Zotb, it == niter - 1 ? b.Zob : nullptr, b, st));
  }
  // CholeskyQR2 on the (bf16-rounded) basis itself
  {
    CopySegments c;
    c.add(b.Ztb, b.Zotb, sizeof(bf16) * q * m);
    c.add(b.Zb, b.Zob, sizeof(bf16) * q * m);
    CB_TRY(copy_if_multi(nullptr, c, st, bt));
  }
  CB_TRY(orthonormalize_b(bt, b.Ztb, b.Zb, m, q, b.Zotb, b.Zob, b, st));
  // B = Zo^T Y (q x n) in both orientations; G = B B^T in fp32
  CB_TRY(g2(bt, n, q, m, b.Ytb, m, b.Zotb, m, nullptr, 0, b.Btb, q, b.Bb, n, nullptr, wd, st, b.tile_counter));
  CB_TRY(g2(bt, q, q, n, b.Bb, n, b.Bb, n, b.G, q, nullptr, 0, nullptr, 0, nullptr, wd, st, b.tile_counter));
  CB_TRY(cholesky_inverse(b.G, (int)q, nullptr, b.status, st, nullptr, bt));
  CB_TRY(jacobi_eigh_from_chol(b.G, (int)q, b.evals, b.V, b.work, b.status != nullptr ? b.status + 1 : nullptr, st, bt));
  CB_TRY(to_bf16(b.V, q, q, q, b.Vb, q, nullptr, 0, nullptr, st, bt));
  // L[m, r] = Zo V_r^T ;  R[r, n] = V_r B (column-scaled by 1 / sqrt(h))
  CB_TRY(g2(bt, m, r, q, b.Zob, q, b.Vb, q, L, r, nullptr, 0, nullptr, 0, nullptr, wd, st, b.tile_counter, aware ? Lsplit : nullptr, 1));
  CB_TRY(g2(bt, r, n, q, b.Vb, q, b.Btb, q, R, n, nullptr, 0, nullptr, 0, aware ? inv_sqrt_h : nullptr, wd, st, b.tile_counter,
            aware ? Rtsplit : nullptr, 2));
  if (!aware) {
    for (int k = 0; k < bt.n; ++k) {
      CB_TRY(scale_cols(at(L, bt, k), m, r, at(b.evals, bt, k), 2, at(L, bt, k), st));
      CB_TRY(scale_rows(at(R, bt, k), r, n, at(b.evals, bt, k), 3, at(R, bt, k), st));
    }
  }
  // rotate the stored basis into its Ritz vectors for the next warm start
  CB_TRY(g2(bt, m, q, q, b.Zob, q, b.Vb, q, nullptr, 0, b.Zb, q, b.Ztb, m, nullptr, wd, st, b.tile_counter));
  CopySegments c;
  c.add(b.Zob, b.Zb, sizeof(bf16) * m * q);
  c.add(b.Zotb, b.Ztb, sizeof(bf16) * m * q);
  return copy_if_multi(nullptr, c, st, bt);
}

// LRbuf = L R for the whole batch (split-bf16 operands, K = 3r, see lr_product_raw).  operands_ready: P.Lb16 / P.Rtb16
// already hold the split operands of the current L, R (written by the epilogues of the contractions that produced them).
static int lr_product_b(const Bt& bt, const LayerPlan& P, int64_t m, int64_t n, int64_t r, cudaStream_t st,
                        bool operands_ready = false) {
  if (!operands_ready) {
    dim3 g1((unsigned)grid_for(m * r, 256 * 4, 1), (unsigned)bt.n);
    split3_kernel<<<g1, 256, 0, st>>>(P.Lcur, m, r, 0, 0, P.Lb16, bt.stride);
    CB_CHECK_LAUNCH();
    dim3 gt((unsigned)((n + 31) / 32), (unsigned)((r + 31) / 32), (unsigned)bt.n);
    split3_t_kernel<<<gt, 256, 0, st>>>(P.Rcur, (int)r, (int)n, 1, P.Rtb16, bt.stride);
    CB_CHECK_LAUNCH();
  }
  return g2(bt, m, n, 3 * r, P.Lb16, 3 * r, P.Rtb16, 3 * r, P.LRbuf, n, nullptr, 0, nullptr, 0, nullptr, P.flags + 4, st, P.tc.tile_counter);
}

// lplr_step_tc for a batch: contractions batched, the fp32 r x r solves and the whole-tensor quantiser per layer
static int lplr_step_b(const Bt& bt, const cb_caldera_params* p, const LayerPlan& P, int64_t m, int64_t n, cudaStream_t st) {
  const int64_t r = p->rank;
  const float* res = p->aware ? P.RES : P.Y;
  int* wd = P.flags + 4;
  auto spd = [&](cudaStream_t s_) -> int {
    CB_TRY(cholesky_inverse(P.Gs, (int)r, P.Linv, P.flags + 2, s_, nullptr, bt));
    return sgemm(r, r, r, 1.f, P.Linv, 1, r, P.Linv, r, 1, P.Ginv, r, 1, false, nullptr, s_, nullptr, bt);   // G^-1 = Linv^T Linv
  };
  // ---- L update (alg.py:163 / :167)
  CB_TRY(to_bf16(P.Rcur, r, n, n, P.Rsb16, n, nullptr, 0, p->aware ? P.sqrt_h : nullptr, st, bt));
  CB_TRY(g2(bt, r, r, n, P.Rsb16, n, P.Rsb16, n, P.Gs, r, nullptr, 0, nullptr, 0, nullptr, wd, st, P.tc.tile_counter));
  CB_TRY(g2(bt, m, r, n, P.tc.Yb, n, P.Rsb16, n, P.Bl, r, nullptr, 0, nullptr, 0, nullptr, wd, st, P.tc.tile_counter));
  CB_TRY(spd(st));
  CB_TRY(sgemm(m, r, r, 1.f, P.Bl, r, 1, P.Ginv, r, 1, P.Ltmp, r, 1, false, nullptr, st, nullptr, bt));
  CB_TRY(quantize_whole_batched(P.Ltmp, m * r, p->l_bits, P.Lcodes_cur, P.Lscale_cur, P.Lcur, P.amax + 1, st, bt));  // alg.py:171-172
  // ---- R update (alg.py:175)
  CB_TRY(to_bf16(P.Lcur, m, r, r, nullptr, 0, P.Ltb16, m, nullptr, st, bt));
  CB_TRY(g2(bt, r, r, m, P.Ltb16, m, P.Ltb16, m, P.Gs, r, nullptr, 0, nullptr, 0, nullptr, wd, st, P.tc.tile_counter));
  CB_TRY(g2(bt, r, n, m, P.Ltb16, m, P.tc.Ytb, m, P.Br, n, nullptr, 0, nullptr, 0, p->aware ? P.inv_sqrt_h : nullptr, wd, st, P.tc.tile_counter));
  CB_TRY(spd(st));
  CB_TRY(sgemm(r, n, r, 1.f, P.Ginv, r, 1, P.Br, n, 1, P.Rtmp, n, 1, false, nullptr, st, nullptr, bt));
  CB_TRY(quantize_whole_batched(P.Rtmp, r * n, p->r_bits, P.Rcodes_cur, P.Rscale_cur, P.Rcur, P.amax + 2, st, bt));  // alg.py:179-180
  // ---- inner error (alg.py:182)
  CB_TRY(lr_product_b(bt, P, m, n, r, st));
  for (int k = 0; k < bt.n; ++k)
    CB_TRY(err_accum(at(res, bt, k), nullptr, 8, nullptr, at(P.LRbuf, bt, k), at(P.w_inner, bt, k), m, n, at(P.dsc, bt, k) + 3, st));
  return CB_OK;
}

static int lplr_refine_b(const Bt& bt, const cb_caldera_params* p, const LayerPlan& P, int64_t m, int64_t n, cudaStream_t st) {
  const int64_t r = p->rank;
  for (int k = 0; k < p->lplr_iters; ++k) {
    CB_TRY(lplr_step_b(bt, p, P, m, n, st));
    CB_TRY(select_inner(P.dsc + 3, P.scalars, P.flags, k == 0, k == p->lplr_iters - 1, st, bt));
    CopySegments best;
    best.add(P.Lb, P.Lcur, sizeof(float) * m * r);
    best.add(P.Rb, P.Rcur, sizeof(float) * r * n);
    best.add(P.Lcodes_in, P.Lcodes_cur, (size_t)m * r * code_bytes(p->l_bits));
    best.add(P.Rcodes_in, P.Rcodes_cur, (size_t)r * n * code_bytes(p->r_bits));
    best.add(P.Lscale_in, P.Lscale_cur, sizeof(float));
    best.add(P.Rscale_in, P.Rscale_cur, sizeof(float));
    CB_TRY(copy_if_multi(P.flags + 1, best, st, bt));
  }
  CopySegments c;
  c.add(P.Lcur, P.Lb, sizeof(float) * m * r);
  c.add(P.Rcur, P.Rb, sizeof(float) * r * n);
  return copy_if_multi(nullptr, c, st, bt);
}

// Which layers the batched driver takes: the production path (tensor-core contractions, identity / diagonal Hessian,
// a sketch the one-CTA-per-layer eigensolver holds).  Everything else runs through cb_caldera_layer.
static bool batch_supported(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind) {
  if (validate_params(p, m, n, h_kind) != CB_OK) return false;
  if (h_kind == CB_H_DENSE || !p->compute_lr || !p->compute_q || !p->use_tensor_cores) return false;
  const int64_t q = default_sketch_width(p, m, n);
  return lowrank_tc_usable(m, n, p->rank, q) && m % 4 == 0 && n % 4 == 0 && q <= 224 && q % 4 == 0;
}

}  // namespace cb

using namespace cb;

extern "C" size_t cb_caldera_layer_workspace_bytes(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind) {
  if (validate_params(p, m, n, h_kind) != CB_OK) return 0;
  Arena a{nullptr, 0, 0};
  LayerPlan L{};
  plan_layer(a, p, m, n, p->scale_w != 0, h_kind, L);
  return a.off + 256;
}

extern "C" int cb_caldera_layer(const cb_caldera_params* p, const float* W, int64_t m, int64_t n, const float* h,
                                int h_kind, const cb_caldera_out* out, void* ws, size_t ws_bytes, void* stream) {
  CB_TRY(validate_params(p, m, n, h_kind));
  PolicyScope policy(p->exec_mode);
  if (W == nullptr || out == nullptr || ws == nullptr) return CB_ERR_ARG;
  if ((h_kind == CB_H_DIAG || h_kind == CB_H_DENSE) && h == nullptr) return CB_ERR_ARG;
  if (out->Q == nullptr || out->L == nullptr || out->R == nullptr || out->errors == nullptr || out->scalars == nullptr)
    return CB_ERR_ARG;
  if (p->compute_q && (out->Q_idxs == nullptr || out->Q_scale == nullptr)) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t r = p->compute_lr ? p->rank : 0;
  const int64_t rr = p->rank;  // L/R outputs always have the parameter rank (alg.py:73-74)
  const int64_t numel = m * n;
  const bool scale_w = p->scale_w != 0;

  Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
  LayerPlan P{};
  CB_TRY(plan_layer(a, p, m, n, scale_w, h_kind, P));
  if (P.quant_factors && (out->L_idxs == nullptr || out->R_idxs == nullptr || out->L_scale == nullptr || out->R_scale == nullptr))
    return CB_ERR_ARG;
  if (p->compute_lr) {
    P.lr.status = P.flags + 2; P.tc.status = P.flags + 2;
    P.lr.seed_dev = out->seed_dev; P.tc.seed_dev = out->seed_dev;
  }

  // ---- initial state: Q = 0, L = 0, R = 0 (alg.py:71-75)
  CB_CUDA(cudaMemsetAsync(P.dsc, 0, sizeof(double) * 4, st));
  CB_CUDA(cudaMemsetAsync(P.flags, 0, sizeof(int) * 8, st));
  CB_CUDA(cudaMemsetAsync(out->L, 0, sizeof(float) * m * rr, st));
  CB_CUDA(cudaMemsetAsync(out->R, 0, sizeof(float) * rr * n, st));
  if (p->compute_q) {
    CB_CUDA(cudaMemsetAsync(out->Q_idxs, 0, (size_t)numel * code_bytes(p->q_bits), st));
    CB_CUDA(cudaMemsetAsync(out->Q_scale, 0, sizeof(float), st));
  }
  if (p->iters * p->n_order > 0) CB_CUDA(cudaMemsetAsync(out->errors, 0, sizeof(float) * p->iters * p->n_order, st));
  if (p->compute_lr) {
    CB_CUDA(cudaMemsetAsync(P.Lcur, 0, sizeof(float) * m * r, st));
    CB_CUDA(cudaMemsetAsync(P.Rcur, 0, sizeof(float) * r * n, st));
  }
  if (P.quant_factors) {
    CB_CUDA(cudaMemsetAsync(P.Lcodes_out, 0, (size_t)m * r * code_bytes(p->l_bits), st));
    CB_CUDA(cudaMemsetAsync(out->R_idxs, 0, (size_t)r * n * code_bytes(p->r_bits), st));
  }

  // ---- global scale, scaled W, error denominator, Hessian vectors
  if (scale_w && !(p->global_scale_in > 0.f)) CB_TRY(sumsq(W, numel, P.dsc + 0, st));
  CB_TRY(finalize_global_scale(P.dsc + 0, numel, p->global_scale_in, scale_w, P.scalars, st));
  CB_TRY(prep_hessian_diag(h_kind == CB_H_DIAG ? h : nullptr, n, p->sigma_reg, p->aware, P.h_eff, P.sqrt_h,
                           P.inv_sqrt_h, P.w_inner, nullptr, st));
  CB_TRY(scale_and_den(W, P.Ws, m, n, P.scalars, P.h_eff, P.dsc + 1, st));
  const float* Ws = scale_w ? P.Ws : W;
  if (P.dense) {
    // alg.py:54 symmetrises only in the activation-aware branch; otherwise H is used as given
    if (p->aware) {
      CB_TRY(symmetrize(h, n, P.Hs, st));
      // alg.py:57-64: lift the spectrum to sigma_reg when its lower end is below it
      CB_TRY(min_eig_shift(P.Hs, n, p->sigma_reg, P.eigwork, nullptr, st));
    } else {
      CB_CUDA(cudaMemcpyAsync(P.Hs, h, sizeof(float) * n * n, cudaMemcpyDeviceToDevice, st));
    }
    CB_CUDA(cudaMemsetAsync(P.dsc + 1, 0, sizeof(double), st));
    CB_TRY(dense_quadratic(P, Ws, nullptr, 8, nullptr, nullptr, m, n, false, P.dsc + 1, st));   // tr(W H W^T)
  }
  if (out->W_scaled != nullptr)
    CB_CUDA(cudaMemcpyAsync(out->W_scaled, Ws, sizeof(float) * numel, cudaMemcpyDeviceToDevice, st));

  // Power iterations of the randomized rank-r step.  The first (cold, random start) step is the
  // inaccurate one: its excess over the exact SVD's residual halves every two iterations (5.9e-4 of
  // the error at 8, 1.7e-4 at 12), while a step warm-started from the previous outer iteration's
  // basis is already below 1e-4 after 3 (oracle/device_model.py study in DESIGN.md section 2).
  const int niter_cold = p->power_iters >= 0 ? p->power_iters : (p->rand_svd ? 2 : 12);
  const int niter_warm = p->power_iters_warm >= 0 ? p->power_iters_warm
                                                  : (p->power_iters >= 0 ? p->power_iters : (p->rand_svd ? 2 : 3));
  bool have_q = false, have_lr = false, lrbuf_valid = false, warm_valid = false, basis_saw_q = false;
  bool y_valid = false;      // P.tc.Yb / Ytb (and RES) already hold the operands of the current Q
  bool amax_valid = false;   // P.amax already holds max |Ws - LRbuf| (computed by the pass that evaluated the error)
  bool updated[8] = {false, false, false, false, false, false, false, false};
  int step = 0;
  for (int it = 0; it < p->iters; ++it) {
    for (int oi = 0; oi < p->n_order; ++oi, ++step) {
      const int which = p->order[oi];
      bool num_ready = false;
      if (which == 0 && p->compute_q) {
        // ---- Q update (maybe_update_Q, alg.py:253-283)
        const float* lrp = nullptr;
        if (p->compute_lr && have_lr) {
          if (!lrbuf_valid) CB_TRY(lr_product(P, m, n, r, st));
          lrbuf_valid = true;
          lrp = P.LRbuf;
        }
        if (!(amax_valid && lrp != nullptr)) CB_TRY(resid_absmax(Ws, lrp, numel, P.amax, st));
        if (P.use_tc && p->compute_lr && m % 4 == 0 && n % 4 == 0) {
          // the rank-r step that follows contracts over Y = (Ws - Q) (.) sqrt(h): build its bf16 operands here
          CB_TRY(quant_form_y_bf16(Ws, lrp, P.h_eff, p->aware ? P.sqrt_h : nullptr, m, n, P.amax, 1e-8f, p->q_bits,
                                   P.codes_cur, P.qscale_cur, P.dsc + 2, P.tc.Yb, P.tc.Ytb,
                                   P.quant_factors ? (p->aware ? P.RES : P.Y) : nullptr, st));
          y_valid = true;
        } else {
          CB_TRY(quant_err(Ws, lrp, P.h_eff, m, n, P.amax, 1e-8f, p->q_bits, P.codes_cur, P.qscale_cur, P.dsc + 2, st));
        }
        have_q = true;
        num_ready = !P.dense;   // the fused numerator is the diagonal metric
        if (P.dense) CB_CUDA(cudaMemsetAsync(P.dsc + 2, 0, sizeof(double), st));
      } else if (which == 1 && p->compute_lr) {
        // ---- LR update (maybe_update_LR / update_LR, alg.py:115-198)
        const void* qcodes = have_q ? P.codes_cur : nullptr;
        const int qbits = p->compute_q ? p->q_bits : 8;
        if (P.use_tc && y_valid) {
          // the Q update that produced the current codes already wrote Yb / Ytb (and the fp32 residual);
          // good for one LR update (a second one in a row rebuilds them below)
          y_valid = false;
        } else if (P.use_tc) {
          // one fused pass: residual -> both bf16 operand orientations (+ fp32 residual for the LPLR loop;
          // when not activation aware that residual is Y itself)
          CB_TRY(form_y_bf16(Ws, qcodes, qbits, P.qscale_cur, p->aware ? P.sqrt_h : nullptr, m, n, P.tc.Yb, P.tc.Ytb,
                             P.quant_factors ? (p->aware ? P.RES : P.Y) : nullptr, st));
        } else {
          CB_TRY(form_y(Ws, qcodes, qbits, P.qscale_cur, (p->aware && !P.dense) ? P.sqrt_h : nullptr, m, n, P.Y, P.RES, st));
        }
        // warm only if the basis being reused was fitted to a residual of the same kind: the step after an
        // LR-first step (Q still zero then) faces a completely different matrix and gets the cold count
        const int niter = (warm_valid && p->warm_start && basis_saw_q) ? niter_warm : niter_cold;
        basis_saw_q = have_q || !p->compute_q;
        if (P.dense && p->aware) {
          // form_y ran with sqrt_h == 1 for the dense case, so P.Y is the plain residual W - Q
          CB_TRY(lowrank_core_dense(P.Y, P.Hs, m, n, r, P.q, niter, p->seed + 0x9E37ull * (uint64_t)step,
                                    warm_valid && p->warm_start, P.Lcur, P.Rcur, P.HP, P.lr, st));
        } else if (P.use_tc) {
          CB_TRY(lowrank_core_tc(m, n, r, P.q, niter, p->seed + 0x9E37ull * (uint64_t)step, p->aware, P.inv_sqrt_h,
                                 warm_valid && p->warm_start, P.Lcur, P.Rcur, nullptr, nullptr, P.tc, st));
        } else {
          CB_TRY(lowrank_core(P.Y, m, n, r, P.q, niter, p->seed + 0x9E37ull * (uint64_t)step, p->aware, P.inv_sqrt_h,
                              warm_valid && p->warm_start, P.Lcur, P.Rcur, P.lr, st));
        }
        warm_valid = true;
        if (P.quant_factors) CB_TRY(lplr_refine(p, P, m, n, P.use_tc, st));
        have_lr = true;
        CB_TRY(lr_product(P, m, n, r, st));
        lrbuf_valid = true;
        amax_valid = false;
      }
      if (!num_ready) {
        const void* qc = have_q ? P.codes_cur : nullptr;
        const float* lrp2 = (have_lr && lrbuf_valid) ? P.LRbuf : nullptr;
        const int qb = p->compute_q ? p->q_bits : 8;
        if (P.dense) CB_TRY(dense_quadratic(P, Ws, qc, qb, P.qscale_cur, lrp2, m, n, false, P.dsc + 2, st));
        else {
          // right after an LR update this pass also delivers the abs-max the next Q update needs
          const bool want_amax = which == 1 && lrp2 != nullptr && p->compute_q;
          CB_TRY(err_accum(Ws, qc, qb, P.qscale_cur, lrp2, P.h_eff, m, n, P.dsc + 2, st, want_amax ? P.amax : nullptr));
          if (want_amax) amax_valid = true;
        }
      }
      updated[oi] = true;
      // `updated` is keyed by name in the reference (alg.py:91), so duplicates in update_order share a flag
      bool all_updated = true;
      for (int k = 0; k < p->n_order; ++k) {
        bool u = false;
        for (int j = 0; j < p->n_order; ++j) u = u || (updated[j] && p->order[j] == p->order[k]);
        all_updated = all_updated && u;
      }
      CB_TRY(select_outer(P.dsc + 2, P.dsc + 1, out->errors, step, P.scalars, P.flags, all_updated ? 1 : 0, st));
      // ---- best_decomp = deepcopy(curr_decomp) (alg.py:107), device side
      CopySegments best;
      if (have_q) {
        best.add(out->Q_idxs, P.codes_cur, (size_t)numel * code_bytes(p->q_bits));
        best.add(out->Q_scale, P.qscale_cur, sizeof(float));
      }
      if (have_lr) {
        best.add(out->L, P.Lcur, sizeof(float) * m * r);
        best.add(out->R, P.Rcur, sizeof(float) * r * n);
        if (P.quant_factors) {
          best.add(P.Lcodes_out, P.Lcodes_in, (size_t)m * r * code_bytes(p->l_bits));
          best.add(out->R_idxs, P.Rcodes_in, (size_t)r * n * code_bytes(p->r_bits));
          best.add(out->L_scale, P.Lscale_in, sizeof(float));
          best.add(out->R_scale, P.Rscale_in, sizeof(float));
        }
      }
      CB_TRY(copy_if_multi(P.flags, best, st));
      // the Q update invalidates nothing; the LR update refreshed LRbuf above
    }
  }

  // ---- materialise the best iterate
  if (p->compute_q) {
    CB_TRY(cb_dequantize_f32(out->Q_idxs, nullptr, out->Q_scale, numel, p->q_bits, 0, out->Q, stream));
    if (out->Q_packed != nullptr) CB_TRY(cb_pack_codes(out->Q_idxs, numel, p->q_bits, out->Q_packed, stream));
  } else {
    CB_CUDA(cudaMemsetAsync(out->Q, 0, sizeof(float) * numel, st));
  }
  if (P.quant_factors) {
    // L_idxs are the codes of L.T flattened (alg.py:171): transpose the (m x r) code matrix
    CB_TRY(transpose_codes(P.Lcodes_out, m, r, code_bytes(p->l_bits), out->L_idxs, st));
    if (out->L_packed != nullptr) CB_TRY(cb_pack_codes(out->L_idxs, m * r, p->l_bits, out->L_packed, stream));
    if (out->R_packed != nullptr) CB_TRY(cb_pack_codes(out->R_idxs, r * n, p->r_bits, out->R_packed, stream));
  }
  // scalars: [0] global_scale [1] min_error [2] best_step; [5..7] int32 bit patterns:
  // cholesky ridge retries, jacobi sweeps of the last solve, tensor-core pipeline watchdog
  CB_CUDA(cudaMemcpyAsync(out->scalars, P.scalars, sizeof(float) * 8, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaMemcpyAsync(out->scalars + 5, P.flags + 2, sizeof(int) * 3, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaMemcpyAsync(out->scalars + 4, P.flags + 5, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

// ---------------------------------------------------------------- batched layer driver
// cb_caldera_layer for `batch` same-shape layers in lock step on one stream.  Layer b reads W + b * stride, h + b *
// stride and writes through out-> pointers + b * stride; its workspace starts at ws + b * stride (stride in bytes, a
// multiple of 256; one slab per layer holds all of these).  Same arithmetic per layer as the single-layer driver with
// the batched contraction engine; results do not depend on the batch size or on a layer's position in the batch.
extern "C" int cb_caldera_batch_supported(const cb_caldera_params* p, int64_t m, int64_t n, int h_kind) {
  return (p != nullptr && batch_supported(p, m, n, h_kind)) ? 1 : 0;
}

extern "C" int cb_caldera_batch(const cb_caldera_params* p, int batch, int64_t stride_bytes, const float* W, int64_t m,
                                int64_t n, const float* h, int h_kind, const cb_caldera_out* out, void* ws, size_t ws_bytes,
                                void* stream) {
  CB_TRY(validate_params(p, m, n, h_kind));
  if (batch < 1 || (batch > 1 && (stride_bytes <= 0 || stride_bytes % 256 != 0))) return CB_ERR_ARG;
  if (!batch_supported(p, m, n, h_kind)) return CB_ERR_UNSUPPORTED;
  if (W == nullptr || out == nullptr || ws == nullptr) return CB_ERR_ARG;
  if (h_kind == CB_H_DIAG && h == nullptr) return CB_ERR_ARG;
  // out->Q (the dense fp32 Q) is optional here: callers that consume the packed codes do not need it
  if (out->L == nullptr || out->R == nullptr || out->errors == nullptr || out->scalars == nullptr ||
      out->Q_idxs == nullptr || out->Q_scale == nullptr)
    return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  Bt bt;
  bt.n = batch; bt.stride = batch > 1 ? stride_bytes : 0;
  const int64_t r = p->rank;
  const int64_t numel = m * n;
  const bool scale_w = p->scale_w != 0;

  Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
  LayerPlan P{};
  CB_TRY(plan_layer(a, p, m, n, scale_w, h_kind, P));
  if (P.quant_factors && (out->L_idxs == nullptr || out->R_idxs == nullptr || out->L_scale == nullptr || out->R_scale == nullptr))
    return CB_ERR_ARG;
  P.lr.status = P.flags + 2; P.tc.status = P.flags + 2;
  P.lr.seed_dev = out->seed_dev; P.tc.seed_dev = out->seed_dev;
  const int qcb = code_bytes(p->q_bits), lcb = code_bytes(p->l_bits), rcb = code_bytes(p->r_bits);

  // ---- initial state: Q = 0, L = 0, R = 0 (alg.py:71-75)
  {
    CopySegments z;
    z.add(P.dsc, nullptr, sizeof(double) * 4);
    z.add(P.flags, nullptr, sizeof(int) * 8);
    z.add(out->L, nullptr, sizeof(float) * m * r);
    z.add(out->R, nullptr, sizeof(float) * r * n);
    z.add(out->Q_idxs, nullptr, (size_t)numel * qcb);
    z.add(out->Q_scale, nullptr, sizeof(float));
    if (p->iters * p->n_order > 0) z.add(out->errors, nullptr, sizeof(float) * p->iters * p->n_order);
    CB_TRY(copy_if_multi(nullptr, z, st, bt));
    CopySegments z2;
    z2.add(P.Lcur, nullptr, sizeof(float) * m * r);
    z2.add(P.Rcur, nullptr, sizeof(float) * r * n);
    z2.add(P.tc.tile_counter, nullptr, sizeof(int) * 4);
    z2.add(P.amax, nullptr, sizeof(float) * 4);
    if (P.quant_factors) {
      z2.add(P.Lcodes_out, nullptr, (size_t)m * r * lcb);
      z2.add(out->R_idxs, nullptr, (size_t)r * n * rcb);
    }
    CB_TRY(copy_if_multi(nullptr, z2, st, bt));
  }
  // ---- global scale, scaled W, error denominator, Hessian vectors
  if (scale_w && !(p->global_scale_in > 0.f))
    for (int b = 0; b < bt.n; ++b) CB_TRY(sumsq(at(W, bt, b), numel, at(P.dsc, bt, b), st));
  CB_TRY(finalize_global_scale(P.dsc + 0, numel, p->global_scale_in, scale_w, P.scalars, st, bt));
  CB_TRY(prep_hessian_diag(h_kind == CB_H_DIAG ? h : nullptr, n, p->sigma_reg, p->aware, P.h_eff, P.sqrt_h, P.inv_sqrt_h,
                           P.w_inner, nullptr, st, bt));
  for (int b = 0; b < bt.n; ++b)   // also delivers max |W / gs|, the abs-max of a first Q update (its residual is W / gs)
    CB_TRY(scale_and_den(at(W, bt, b), at(P.Ws, bt, b), m, n, at(P.scalars, bt, b), at(P.h_eff, bt, b), at(P.dsc, bt, b) + 1, st,
                         at(P.amax, bt, b)));
  const float* Ws = scale_w ? P.Ws : W;
  if (out->W_scaled != nullptr) CB_TRY(copy_b(bt, out->W_scaled, Ws, sizeof(float) * numel, st));

  const int niter_cold = p->power_iters >= 0 ? p->power_iters : (p->rand_svd ? 2 : 12);
  const int niter_warm = p->power_iters_warm >= 0 ? p->power_iters_warm
                                                  : (p->power_iters >= 0 ? p->power_iters : (p->rand_svd ? 2 : 3));
  bool have_q = false, have_lr = false, lrbuf_valid = false, warm_valid = false, basis_saw_q = false;
  bool y_valid = false;
  bool amax_valid = true;                  // P.amax holds max |Ws - L R| of the current L, R (L = R = 0 so far: max |Ws|)
  bool updated[8] = {false, false, false, false, false, false, false, false};
  float* res_out = P.quant_factors ? (p->aware ? P.RES : P.Y) : nullptr;   // fp32 residual W - Q for the LPLR loop
  int step = 0;
  for (int it = 0; it < p->iters; ++it) {
    for (int oi = 0; oi < p->n_order; ++oi, ++step) {
      const int which = p->order[oi];
      bool num_ready = false;
      if (which == 0) {
        // ---- Q update (maybe_update_Q, alg.py:253-283) fused with the bf16 operand builder of the next rank-r step
        {
          const float* lrp = nullptr;
          if (have_lr) {
            if (!lrbuf_valid) CB_TRY(lr_product_b(bt, P, m, n, r, st));
            lrbuf_valid = true;
            lrp = P.LRbuf;
          }
          for (int b = 0; b < bt.n; ++b) {
            if (!amax_valid) CB_TRY(resid_absmax(at(Ws, bt, b), at(lrp, bt, b), numel, at(P.amax, bt, b), st));
            CB_TRY(quant_form_y_bf16(at(Ws, bt, b), at(lrp, bt, b), at(P.h_eff, bt, b), p->aware ? at(P.sqrt_h, bt, b) : nullptr, m, n,
                                     at(P.amax, bt, b), 1e-8f, p->q_bits, at(P.codes_cur, bt, b), at(P.qscale_cur, bt, b),
                                     at(P.dsc, bt, b) + 2, at(P.tc.Yb, bt, b), at(P.tc.Ytb, bt, b), at(res_out, bt, b), st));
          }
        }
        y_valid = true;
        have_q = true;
        num_ready = true;
      } else {
        // ---- LR update (maybe_update_LR / update_LR, alg.py:115-198)
        if (y_valid) {
          y_valid = false;       // the Q update that produced the current codes already wrote Yb / Ytb (and the residual)
        } else {
          for (int b = 0; b < bt.n; ++b)
            CB_TRY(form_y_bf16(at(Ws, bt, b), have_q ? at(P.codes_cur, bt, b) : nullptr, p->q_bits, at(P.qscale_cur, bt, b),
                               p->aware ? at(P.sqrt_h, bt, b) : nullptr, m, n, at(P.tc.Yb, bt, b), at(P.tc.Ytb, bt, b),
                               at(res_out, bt, b), st));
        }
        const int niter = (warm_valid && p->warm_start && basis_saw_q) ? niter_warm : niter_cold;
        basis_saw_q = have_q;
        // 16-bit factors: L and R leave the rank-r step as they are, so their split-bf16 operands come out of the
        // epilogues that produce them
        const bool split_in_epilogue = p->aware && !P.quant_factors;
        CB_TRY(lowrank_core_b(bt, m, n, r, P.q, niter, p->seed + 0x9E37ull * (uint64_t)step, p->aware, P.inv_sqrt_h,
                              warm_valid && p->warm_start, P.Lcur, P.Rcur, P.tc, st, split_in_epilogue ? P.Lb16 : nullptr,
                              split_in_epilogue ? P.Rtb16 : nullptr));
        warm_valid = true;
        if (P.quant_factors) CB_TRY(lplr_refine_b(bt, p, P, m, n, st));
        have_lr = true;
        amax_valid = false;
        CB_TRY(lr_product_b(bt, P, m, n, r, st, split_in_epilogue));
        lrbuf_valid = true;
      }
      if (!num_ready) {
        // right after an LR update this pass also delivers the abs-max the next Q update needs
        const bool want_amax = which == 1;
        for (int b = 0; b < bt.n; ++b)
          CB_TRY(err_accum(at(Ws, bt, b), have_q ? at(P.codes_cur, bt, b) : nullptr, p->q_bits, at(P.qscale_cur, bt, b),
                           (have_lr && lrbuf_valid) ? at(P.LRbuf, bt, b) : nullptr, at(P.h_eff, bt, b), m, n, at(P.dsc, bt, b) + 2, st,
                           want_amax ? at(P.amax, bt, b) : nullptr));
        if (want_amax) amax_valid = true;
      }
      updated[oi] = true;
      bool all_updated = true;
      for (int k = 0; k < p->n_order; ++k) {
        bool u = false;
        for (int j = 0; j < p->n_order; ++j) u = u || (updated[j] && p->order[j] == p->order[k]);
        all_updated = all_updated && u;
      }
      CB_TRY(select_outer(P.dsc + 2, P.dsc + 1, out->errors, step, P.scalars, P.flags, all_updated ? 1 : 0, st, bt));
      // ---- best_decomp = deepcopy(curr_decomp) (alg.py:107), device side
      CopySegments best;
      if (have_q) {
        best.add(out->Q_idxs, P.codes_cur, (size_t)numel * qcb);
        best.add(out->Q_scale, P.qscale_cur, sizeof(float));
      }
      if (have_lr) {
        best.add(out->L, P.Lcur, sizeof(float) * m * r);
        best.add(out->R, P.Rcur, sizeof(float) * r * n);
        if (P.quant_factors) {
          best.add(P.Lcodes_out, P.Lcodes_in, (size_t)m * r * lcb);
          best.add(out->R_idxs, P.Rcodes_in, (size_t)r * n * rcb);
          best.add(out->L_scale, P.Lscale_in, sizeof(float));
          best.add(out->R_scale, P.Rscale_in, sizeof(float));
        }
      }
      CB_TRY(copy_if_multi(P.flags, best, st, bt));
    }
  }
  // ---- materialise the best iterate
  for (int b = 0; b < bt.n; ++b) {
    if (out->Q != nullptr)
      CB_TRY(cb_dequantize_f32(at(out->Q_idxs, bt, b), nullptr, at(out->Q_scale, bt, b), numel, p->q_bits, 0, at(out->Q, bt, b), stream));
    if (out->Q_packed != nullptr) CB_TRY(cb_pack_codes(at(out->Q_idxs, bt, b), numel, p->q_bits, at(out->Q_packed, bt, b), stream));
    if (P.quant_factors) {
      CB_TRY(transpose_codes(at(P.Lcodes_out, bt, b), m, r, lcb, at(out->L_idxs, bt, b), st));
      if (out->L_packed != nullptr) CB_TRY(cb_pack_codes(at(out->L_idxs, bt, b), m * r, p->l_bits, at(out->L_packed, bt, b), stream));
      if (out->R_packed != nullptr) CB_TRY(cb_pack_codes(at(out->R_idxs, bt, b), r * n, p->r_bits, at(out->R_packed, bt, b), stream));
    }
  }
  CopySegments fin;
  fin.add(out->scalars, P.scalars, sizeof(float) * 4);
  fin.add(out->scalars + 5, P.flags + 2, sizeof(int) * 3);
  fin.add(out->scalars + 4, P.flags + 5, sizeof(int));
  return copy_if_multi(nullptr, fin, st, bt);
}

// ---------------------------------------------------------------- Hessian accumulation
// H += X^T X for a batch X (T x n) of calibration activations -- the step main.py:307-311 runs on the CPU in
// fp64, one sample at a time -- and/or its diagonal sum_t X[t, j]^2.  The product runs on the tcgen05
// kernel with split-bf16 operands ([hi | hi | lo] x [hi | lo | hi], K = 3T: ~16 mantissa bits per factor).
__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) y[i] += x[i];
}
__global__ void __launch_bounds__(256)
colsumsq_kernel(const float* __restrict__ X, int64_t T, int64_t n, int64_t ldx, float* __restrict__ hdiag) {
  // one thread per column, rows split over blockIdx.y; fp32 partial sums added with one atomic per thread
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t rows_per = (T + gridDim.y - 1) / gridDim.y;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per, t1 = t0 + rows_per < T ? t0 + rows_per : T;
  float acc = 0.f;
  for (int64_t t = t0; t < t1; ++t) { const float v = X[t * ldx + j]; acc = fmaf(v, v, acc); }
  if (t1 > t0) atomicAdd(hdiag + j, acc);
}

extern "C" size_t cb_hessian_accumulate_workspace_bytes(int64_t T, int64_t n) {
  if (T <= 0 || n <= 0) return 0;
  const int64_t T8 = T / 8 * 8;
  return (size_t)2 * n * 3 * T8 * sizeof(bf16) + (size_t)n * n * sizeof(float) + kSplitWsBytes + 4 * 256;
}

extern "C" int cb_hessian_accumulate_f32(const float* X, int64_t T, int64_t n, float* H, float* hdiag, int* error_flag,
                                         void* ws, size_t ws_bytes, void* stream) {
  if (X == nullptr || T <= 0 || n <= 0 || (H == nullptr && hdiag == nullptr)) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (hdiag != nullptr) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(T >= 1024 ? 32 : (T >= 64 ? 8 : 1)));
    colsumsq_kernel<<<grid, 256, 0, st>>>(X, T, n, n, hdiag);
    CB_CHECK_LAUNCH();
  }
  if (H == nullptr) return CB_OK;
  const int64_t T8 = T / 8 * 8;
  if (n % 8 == 0 && T8 > 0 && aligned16(X)) {
    if (ws == nullptr || ws_bytes < cb_hessian_accumulate_workspace_bytes(T, n)) return CB_ERR_WORKSPACE;
    Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
    bf16* A3 = a.take<bf16>(n * 3 * T8);
    bf16* B3 = a.take<bf16>(n * 3 * T8);
    float* tmp = a.take<float>(n * n);
    SplitWs sw;
    sw.buf = a.take<float>(kSplitWsBytes / sizeof(float)); sw.bytes = kSplitWsBytes;
    if (!a.ok()) return CB_ERR_WORKSPACE;
    split3_kernel<<<grid_for(T8 * n, 256 * 4, 4), 256, 0, st>>>(X, T8, n, 1, 0, A3);   // rows = features: [hi | hi | lo]
    CB_CHECK_LAUNCH();
    split3_kernel<<<grid_for(T8 * n, 256 * 4, 4), 256, 0, st>>>(X, T8, n, 1, 1, B3);   //                  [hi | lo | hi]
    CB_CHECK_LAUNCH();
    CB_TRY(gemm_tc(n, n, 3 * T8, 1.f, A3, 3 * T8, B3, 3 * T8, tmp, n, nullptr, 0, nullptr, 0, nullptr, nullptr, 0,
                   error_flag, nullptr, st, &sw));
    axpy_kernel<<<grid_for(n * n, 256 * 4, 4), 256, 0, st>>>(tmp, H, n * n);
    CB_CHECK_LAUNCH();
    if (T8 < T)   // the last T % 8 rows: fp32 SIMT, accumulated in place
      CB_TRY(sgemm(n, n, T - T8, 1.f, X + T8 * n, 1, n, X + T8 * n, n, 1, H, n, 1, true, nullptr, st));
    return CB_OK;
  }
  // ragged shapes: fp32 SIMT contraction, accumulated in place
  return sgemm(n, n, T, 1.f, X, 1, n, X, n, 1, H, n, 1, true, nullptr, st);
}

// ---------------------------------------------------------------- stand-alone stages (C ABI)
extern "C" size_t cb_lowrank_init_workspace_bytes(int64_t m, int64_t n, int64_t r, int64_t q_width, int h_kind) {
  if (m <= 0 || n <= 0 || r <= 0 || q_width < r || q_width > 512) return 0;
  (void)h_kind;
  Arena a{nullptr, 0, 0};
  a.take<float>(4 * n);   // h vectors
  a.take<float>(m * n);   // Y
  a.take<int>(8);
  plan_lowrank(a, m, n, q_width, nullptr, nullptr);
  return a.off + 256;
}

extern "C" int cb_lowrank_init(const float* A, int64_t m, int64_t n, const float* h, int h_kind, int64_t r,
                               int64_t q_width, int niter, uint64_t seed, int aware, float* L, float* R,
                               float* sigma, void* ws, size_t ws_bytes, void* stream) {
  if (A == nullptr || L == nullptr || R == nullptr || ws == nullptr || m <= 0 || n <= 0) return CB_ERR_ARG;
  if (r < 1 || r > m || r > n || q_width < r || q_width > m || q_width > n || niter < 0) return CB_ERR_ARG;
  if (q_width > 512) return CB_ERR_UNSUPPORTED;
  if (h_kind == CB_H_DENSE) return CB_ERR_UNSUPPORTED;
  if (h_kind == CB_H_DIAG && h == nullptr) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
  float* hv = a.take<float>(4 * n);
  float* Y = a.take<float>(m * n);
  int* status = a.take<int>(8);
  LowrankBufs b = plan_lowrank(a, m, n, q_width, nullptr, status);
  if (!a.ok()) return CB_ERR_WORKSPACE;
  CB_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * 8, st));
  CB_TRY(prep_hessian_diag(h_kind == CB_H_DIAG ? h : nullptr, n, 0.f, aware, hv, hv + n, hv + 2 * n, hv + 3 * n, nullptr, st));
  CB_TRY(form_y(A, nullptr, 8, nullptr, aware ? hv + n : nullptr, m, n, Y, nullptr, st));
  CB_TRY(lowrank_core(Y, m, n, r, q_width, niter, seed, aware, hv + 2 * n, false, L, R, b, st));
  if (sigma != nullptr) {
    sqrt_vec_kernel<<<1, 256, 0, st>>>(b.evals, (int)r, sigma);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

extern "C" size_t cb_weighted_error_workspace_bytes(int64_t m, int64_t n, int64_t r, int h_kind) {
  (void)h_kind;
  if (m <= 0 || n <= 0) return 0;
  return (r > 0 ? sizeof(float) * (size_t)m * n : 0) + 1024;
}

extern "C" int cb_weighted_error(const float* W, int64_t m, int64_t n, const void* q_codes, int q_bits,
                                 const float* q_scale, const float* L, const float* R, int64_t r, const float* h,
                                 int h_kind, double* out_num, double* out_den, void* ws, size_t ws_bytes,
                                 void* stream) {
  if (W == nullptr || out_num == nullptr || m <= 0 || n <= 0) return CB_ERR_ARG;
  if (h_kind == CB_H_DENSE) return CB_ERR_UNSUPPORTED;
  if (h_kind == CB_H_DIAG && h == nullptr) return CB_ERR_ARG;
  if (q_codes != nullptr && (!bits_ok(q_bits) || q_scale == nullptr)) return CB_ERR_BITS;
  cudaStream_t st = (cudaStream_t)stream;
  const float* hw = h_kind == CB_H_DIAG ? h : nullptr;
  float* LRbuf = nullptr;
  if (L != nullptr && R != nullptr && r > 0) {
    if (ws == nullptr || ws_bytes < sizeof(float) * (size_t)m * n) return CB_ERR_WORKSPACE;
    LRbuf = reinterpret_cast<float*>(ws);
    CB_TRY(sgemm(m, n, r, 1.f, L, r, 1, R, n, 1, LRbuf, n, 1, false, nullptr, st));
  }
  CB_TRY(err_accum(W, q_codes, q_codes != nullptr ? q_bits : 8, q_scale, LRbuf, hw, m, n, out_num, st));
  if (out_den != nullptr) CB_TRY(err_accum(W, nullptr, 8, nullptr, nullptr, hw, m, n, out_den, st));
  return CB_OK;
}

#ifdef CB_MEASURE
extern "C" int cb_probe_err_pass(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR,
                                 const float* w, int64_t m, int64_t n, double* num, float* amax_next, void* stream) {
  if (Ws == nullptr || num == nullptr || m <= 0 || n <= 0) return CB_ERR_ARG;
  return cb::err_accum(Ws, codes, bits, qscale, LR, w, m, n, num, (cudaStream_t)stream, amax_next);
}
#endif

// ---------------------------------------------------------------- one LPLR iteration as a stage (C ABI)
namespace cb {
struct LplrStagePlan {
  LayerPlan P;
  float* res_w;      // aware: res (.) sqrt(h) in fp32 is not needed; bf16 operands are built straight from res
  void* Lcodes_t;    // L codes in the reference's order ((L^T).flatten(), alg.py:171)
};
static int plan_lplr_stage(Arena& a, int64_t m, int64_t n, int64_t r, int l_bits, int r_bits, bool use_tc, LplrStagePlan& S) {
  LayerPlan& L = S.P;
  L.dense = false; L.quant_factors = true; L.use_tc = use_tc; L.q = 0;
  L.dsc = a.take<double>(4);
  L.flags = a.take<int>(8);
  L.scalars = a.take<float>(8);
  L.h_eff = a.take<float>(n); L.sqrt_h = a.take<float>(n); L.inv_sqrt_h = a.take<float>(n); L.w_inner = a.take<float>(n);
  L.LRbuf = a.take<float>(m * n);
  L.Lcur = a.take<float>(m * r); L.Rcur = a.take<float>(r * n);
  L.Rw = a.take<float>(r * n); L.Gs = a.take<float>(r * r); L.Linv = a.take<float>(r * r); L.Ginv = a.take<float>(r * r);
  L.Bl = a.take<float>(m * r); L.Br = a.take<float>(r * n); L.Ltmp = a.take<float>(m * r); L.Rtmp = a.take<float>(r * n);
  L.Lcodes_cur = a.take<uint8_t>(m * r * code_bytes(l_bits));
  L.Rcodes_cur = a.take<uint8_t>(r * n * code_bytes(r_bits));
  L.Lscale_cur = a.take<float>(4); L.Rscale_cur = a.take<float>(4);
  L.lr.sw.buf = a.take<float>(kSplitWsBytes / sizeof(float)); L.lr.sw.bytes = kSplitWsBytes;
  L.tc.sw = L.lr.sw;
  if (use_tc) {
    L.tc.Yb = a.take<bf16>(m * n); L.tc.Ytb = a.take<bf16>(n * m);
    L.Lb16 = a.take<bf16>(3 * m * r); L.Rtb16 = a.take<bf16>(3 * n * r);
    L.Rsb16 = a.take<bf16>(r * n); L.Ltb16 = a.take<bf16>(r * m);
  }
  S.Lcodes_t = a.take<uint8_t>(m * r * code_bytes(l_bits));
  return a.ok() ? CB_OK : CB_ERR_WORKSPACE;
}
}  // namespace cb

extern "C" size_t cb_lplr_iter_workspace_bytes(int64_t m, int64_t n, int64_t r, int l_bits, int r_bits, int use_tensor_cores) {
  if (m <= 0 || n <= 0 || r < 1 || r > m || r > n || r > 512 || !bits_ok(l_bits) || !bits_ok(r_bits)) return 0;
  Arena a{nullptr, 0, 0};
  LplrStagePlan S{};
  plan_lplr_stage(a, m, n, r, l_bits, r_bits, use_tensor_cores != 0 && lowrank_tc_usable(m, n, r, 16), S);
  return a.off + 256;
}

extern "C" int cb_lplr_iter(const float* res, int64_t m, int64_t n, const float* h, int h_kind, int aware, int64_t r,
                            int l_bits, int r_bits, const float* R_in, float* L_pre, void* L_idxs, float* L_scale,
                            float* L_hat, float* R_pre, void* R_idxs, float* R_scale, float* R_hat, double* err_sq,
                            int use_tensor_cores, int* status, void* ws, size_t ws_bytes, void* stream) {
  if (res == nullptr || R_in == nullptr || ws == nullptr || m <= 0 || n <= 0) return CB_ERR_ARG;
  if (r < 1 || r > m || r > n) return CB_ERR_ARG;
  if (r > 512) return CB_ERR_UNSUPPORTED;
  if (!bits_ok(l_bits) || !bits_ok(r_bits)) return CB_ERR_BITS;
  if (h_kind == CB_H_DENSE) return CB_ERR_UNSUPPORTED;
  if (h_kind != CB_H_IDENTITY && h_kind != CB_H_DIAG) return CB_ERR_ARG;
  if (h_kind == CB_H_DIAG && h == nullptr) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool use_tc = use_tensor_cores != 0 && lowrank_tc_usable(m, n, r, 16);
  Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
  LplrStagePlan S{};
  CB_TRY(plan_lplr_stage(a, m, n, r, l_bits, r_bits, use_tc, S));
  LayerPlan& P = S.P;
  cb_caldera_params prm{};
  prm.rank = (int32_t)r; prm.l_bits = l_bits; prm.r_bits = r_bits; prm.aware = aware != 0; prm.lplr_iters = 1;
  CB_CUDA(cudaMemsetAsync(P.dsc, 0, sizeof(double) * 4, st));
  CB_CUDA(cudaMemsetAsync(P.flags, 0, sizeof(int) * 8, st));
  CB_TRY(prep_hessian_diag(h_kind == CB_H_DIAG ? h : nullptr, n, 0.f, prm.aware, P.h_eff, P.sqrt_h, P.inv_sqrt_h, P.w_inner,
                           nullptr, st));
  // lplr_step reads the unweighted residual through P.RES (aware) / P.Y (not aware)
  P.RES = const_cast<float*>(res); P.Y = const_cast<float*>(res);
  if (use_tc) CB_TRY(to_bf16(res, m, n, n, P.tc.Yb, n, P.tc.Ytb, m, prm.aware ? P.sqrt_h : nullptr, st));
  CB_CUDA(cudaMemcpyAsync(P.Rcur, R_in, sizeof(float) * r * n, cudaMemcpyDeviceToDevice, st));
  CB_TRY(use_tc ? lplr_step_tc(&prm, P, m, n, st) : lplr_step(&prm, P, m, n, st));
  if (L_pre != nullptr) CB_CUDA(cudaMemcpyAsync(L_pre, P.Ltmp, sizeof(float) * m * r, cudaMemcpyDeviceToDevice, st));
  if (R_pre != nullptr) CB_CUDA(cudaMemcpyAsync(R_pre, P.Rtmp, sizeof(float) * r * n, cudaMemcpyDeviceToDevice, st));
  if (L_hat != nullptr) CB_CUDA(cudaMemcpyAsync(L_hat, P.Lcur, sizeof(float) * m * r, cudaMemcpyDeviceToDevice, st));
  if (R_hat != nullptr) CB_CUDA(cudaMemcpyAsync(R_hat, P.Rcur, sizeof(float) * r * n, cudaMemcpyDeviceToDevice, st));
  if (L_scale != nullptr) CB_CUDA(cudaMemcpyAsync(L_scale, P.Lscale_cur, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (R_scale != nullptr) CB_CUDA(cudaMemcpyAsync(R_scale, P.Rscale_cur, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (L_idxs != nullptr) CB_TRY(transpose_codes(P.Lcodes_cur, m, r, code_bytes(l_bits), L_idxs, st));
  if (R_idxs != nullptr)
    CB_CUDA(cudaMemcpyAsync(R_idxs, P.Rcodes_cur, (size_t)r * n * code_bytes(r_bits), cudaMemcpyDeviceToDevice, st));
  if (err_sq != nullptr) CB_CUDA(cudaMemcpyAsync(err_sq, P.dsc + 3, sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (status != nullptr) CB_CUDA(cudaMemcpyAsync(status, P.flags + 2, sizeof(int) * 3, cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

// ---------------------------------------------------------------- Convex-CALDERA prox solver
// Accelerated proximal gradient on the reduced program of oracle/convex_oracle.py
// (solve_convex_optimization, RCR/convex_caldera/decomposition/convex_caldera.py:128-241, as
// documented in README.md:89-93).  One iteration = one fused pass forming both prox arguments,
// a rank-capped randomized singular-value thresholding of V_L (the same subspace iteration /
// Rayleigh-Ritz machinery as LR_init, warm-started from the previous iterate), a dense rebuild of
// L, and a fused radial shrink of R with the smooth part of the objective.
namespace cb {
struct ConvexPlan {
  float *VL, *VR, *Lw;
  bf16 *Lb16, *Rtb16;
  int* flags;
  LowrankBufs lr;
  LowrankTcBufs tc;
  bool use_tc;
};
static int plan_convex(Arena& a, int64_t m, int64_t n, int64_t r, int64_t q, int use_tensor_cores, ConvexPlan& P) {
  P.VL = a.take<float>(m * n);
  P.VR = a.take<float>(m * n);
  P.Lw = a.take<float>(m * r);
  P.flags = a.take<int>(8);
  P.lr = plan_lowrank(a, m, n, q, nullptr, nullptr);
  P.use_tc = use_tensor_cores != 0 && lowrank_tc_usable(m, n, r, q);
  if (P.use_tc) {
    P.tc = plan_lowrank_tc(a, m, n, q, nullptr, P.lr.sw);
    P.Lb16 = a.take<bf16>(3 * m * r);
    P.Rtb16 = a.take<bf16>(3 * n * r);
  }
  return a.ok() ? CB_OK : CB_ERR_WORKSPACE;
}
}  // namespace cb

extern "C" size_t cb_convex_prox_workspace_bytes(int64_t m, int64_t n, int64_t rank_cap, int64_t q_width,
                                                 int use_tensor_cores) {
  if (m <= 0 || n <= 0 || rank_cap < 1 || q_width < rank_cap || q_width > 512 || q_width > m || q_width > n) return 0;
  Arena a{nullptr, 0, 0};
  ConvexPlan P{};
  plan_convex(a, m, n, rank_cap, q_width, use_tensor_cores, P);
  return a.off + 256;
}

// h: diagonal of the Hessian (nullptr = identity) or, with `dense`, the symmetric n x n matrix itself.  For a dense
// Hessian the gradient (W - Y_L - Y_R) H is one fp32 contraction per iteration and the smooth term of the objective
// (one more contraction) is evaluated for the last iterate of the call only -- the caller reads it back once per call.
static int convex_prox_iters(const float* W, const float* h, bool dense, int64_t m, int64_t n, float mu, float tau_star,
                             float lambda_reg, float kappa, float q0, float step_t, int64_t rank_cap,
                             int64_t q_width, int power_iters, uint64_t seed, int warm, int use_tensor_cores,
                             int n_iters, double* theta_io, float* L, float* Lp, float* R, float* Rp,
                             float* Lf, float* Rf, float* svals, double* scalars, void* ws, size_t ws_bytes,
                             void* stream) {
  if (dense && h == nullptr) return CB_ERR_ARG;
  if (W == nullptr || L == nullptr || Lp == nullptr || R == nullptr || Rp == nullptr || Lf == nullptr ||
      Rf == nullptr || svals == nullptr || scalars == nullptr || theta_io == nullptr || ws == nullptr)
    return CB_ERR_ARG;
  if (m <= 0 || n <= 0 || rank_cap < 1 || q_width < rank_cap || q_width > m || q_width > n || n_iters < 0) return CB_ERR_ARG;
  if (q_width > 512) return CB_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t r = rank_cap, q = q_width;
  Arena a{reinterpret_cast<uint8_t*>(ws), 0, ws_bytes};
  ConvexPlan P{};
  CB_TRY(plan_convex(a, m, n, r, q, use_tensor_cores, P));
  P.lr.status = P.flags;
  P.tc.status = P.flags;
  if (!warm) CB_CUDA(cudaMemsetAsync(P.flags, 0, sizeof(int) * 8, st));
  const bool constrained = mu < 0.f;
  float *Lc = L, *Lprev = Lp, *Rc = R, *Rprev = Rp;
  double theta = *theta_io;
  bool warm_valid = warm != 0;
  for (int it = 0; it < n_iters; ++it) {
    const double theta_next = (1.0 + sqrt(1.0 + 4.0 * theta * theta)) / 2.0;
    const float beta = (float)((theta - 1.0) / theta_next);
    theta = theta_next;
    CB_CUDA(cudaMemsetAsync(scalars + 3, 0, sizeof(double) * 2, st));        // smooth term, ||V_R||^2
    if (dense) {
      CB_TRY(cvx_resid(W, Lc, Lprev, Rc, Rprev, m, n, beta, P.VR, st));
      CB_TRY(sgemm(m, n, n, 1.f, P.VR, n, 1, h, n, 1, P.VL, n, 1, false, nullptr, st));
      CB_TRY(cvx_point_dense(Lc, Lprev, Rc, Rprev, m, n, beta, step_t, P.VL, P.VR, scalars + 4, st));
    } else {
      CB_TRY(cvx_point(W, Lc, Lprev, Rc, Rprev, h, m, n, beta, step_t, P.VL, P.VR, scalars + 4, st));
    }
    // singular-value thresholding of V_L: top-r factors U sqrt(S), sqrt(S) V^T and S^2
    const float* sig2;
    if (P.use_tc) {
      CB_TRY(to_bf16(P.VL, m, n, n, P.tc.Yb, n, P.tc.Ytb, m, nullptr, st));
      CB_TRY(lowrank_core_tc(m, n, r, q, power_iters, seed + 977ull * (uint64_t)it, 0, nullptr, warm_valid, Lf, Rf,
                             nullptr, nullptr, P.tc, st));
      sig2 = P.tc.evals;
    } else {
      CB_TRY(lowrank_core(P.VL, m, n, r, q, power_iters, seed + 977ull * (uint64_t)it, 0, nullptr, warm_valid, Lf, Rf,
                          P.lr, st));
      sig2 = P.lr.evals;
    }
    warm_valid = true;
    CB_TRY(cvx_shrink(sig2, (int)r, step_t * (constrained ? 0.f : mu), tau_star, constrained ? 1 : 0, scalars + 4,
                      step_t * lambda_reg, kappa, q0, svals + r, svals, scalars, st));
    CB_TRY(scale_cols(Lf, m, r, svals + r, 0, P.Lw, st));
    // the new L overwrites the buffer of the iterate before last, then roles rotate
    CB_TRY(lr_product_raw(P.use_tc, P.Lw, Rf, m, n, r, Lprev, P.Lb16, P.Rtb16, P.flags + 2, st));
    if (dense) {
      const bool last = it == n_iters - 1;
      CB_TRY(cvx_finish_dense(W, Lprev, P.VR, m, n, scalars, Rprev, last ? P.VL : nullptr, st));
      if (last) {                       // smooth term 1/2 sum (E H) (.) E of the iterate the caller will read
        CB_TRY(sgemm(m, n, n, 1.f, P.VL, n, 1, h, n, 1, P.VR, n, 1, false, nullptr, st));
        CB_TRY(dot_accum(P.VR, P.VL, m * n, scalars + 3, st));
        CB_TRY(scale_double(scalars + 3, 0.5, st));
      }
    } else {
      CB_TRY(cvx_finish(W, Lprev, P.VR, h, m, n, scalars, Rprev, scalars + 3, st));
    }
    float* tmp = Lc; Lc = Lprev; Lprev = tmp;
    tmp = Rc; Rc = Rprev; Rprev = tmp;
  }
  if (Lc != L) {
    // odd number of iterations: swap contents so that (L, R) hold the iterate and (Lp, Rp) the previous one
    CB_CUDA(cudaMemcpyAsync(P.VL, L, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaMemcpyAsync(L, Lp, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaMemcpyAsync(Lp, P.VL, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaMemcpyAsync(P.VR, R, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaMemcpyAsync(R, Rp, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaMemcpyAsync(Rp, P.VR, sizeof(float) * m * n, cudaMemcpyDeviceToDevice, st));
  }
  *theta_io = theta;
  return CB_OK;
}

extern "C" int cb_convex_prox_iters(const float* W, const float* h, int64_t m, int64_t n, float mu, float tau_star,
                                    float lambda_reg, float kappa, float q0, float step_t, int64_t rank_cap,
                                    int64_t q_width, int power_iters, uint64_t seed, int warm, int use_tensor_cores,
                                    int n_iters, double* theta_io, float* L, float* Lp, float* R, float* Rp,
                                    float* Lf, float* Rf, float* svals, double* scalars, void* ws, size_t ws_bytes,
                                    void* stream) {
  return convex_prox_iters(W, h, false, m, n, mu, tau_star, lambda_reg, kappa, q0, step_t, rank_cap, q_width, power_iters,
                           seed, warm, use_tensor_cores, n_iters, theta_io, L, Lp, R, Rp, Lf, Rf, svals, scalars, ws,
                           ws_bytes, stream);
}

extern "C" int cb_convex_prox_iters_dense(const float* W, const float* Hs, int64_t m, int64_t n, float mu, float tau_star,
                                          float lambda_reg, float kappa, float q0, float step_t, int64_t rank_cap,
                                          int64_t q_width, int power_iters, uint64_t seed, int warm, int use_tensor_cores,
                                          int n_iters, double* theta_io, float* L, float* Lp, float* R, float* Rp,
                                          float* Lf, float* Rf, float* svals, double* scalars, void* ws, size_t ws_bytes,
                                          void* stream) {
  return convex_prox_iters(W, Hs, true, m, n, mu, tau_star, lambda_reg, kappa, q0, step_t, rank_cap, q_width, power_iters,
                           seed, warm, use_tensor_cores, n_iters, theta_io, L, Lp, R, Rp, Lf, Rf, svals, scalars, ws,
                           ws_bytes, stream);
}

extern "C" int cb_scale_f32(const float* X, int64_t rows, int64_t cols, const float* v, int axis, int mode, float* out,
                            void* stream) {
  if (X == nullptr || v == nullptr || out == nullptr || rows <= 0 || cols <= 0 || mode < 0 || mode > 3) return CB_ERR_ARG;
  return axis == 0 ? cb::scale_rows(X, rows, cols, v, mode, out, (cudaStream_t)stream)
                   : cb::scale_cols(X, rows, cols, v, mode, out, (cudaStream_t)stream);
}

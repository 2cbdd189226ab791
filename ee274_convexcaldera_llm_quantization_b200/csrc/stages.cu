// Fused element-wise stages of the CALDERA outer loop.
//
// Every kernel here makes exactly one pass over an m x n operand with 128-bit accesses and
// folds in everything the reference does in separate torch ops around it:
//   scale_and_den    W / global_scale           + denominator  sum_j h_j W_ij^2   (alg.py:42, 298)
//   resid_absmax     residual = W - L R         + abs-max for the whole-tensor scale (alg.py:262, quantization.py:262)
//   quant_err        quantise + dequantise      + error numerator sum_j h_j (res - Q)^2 (alg.py:280-283, 298)
//   form_y           residual = W - Q           + column weighting by sqrt(h)  (alg.py:124, 211)
//   err_accum        E = W - Q - L R            + sum_j w_j E_ij^2              (alg.py:182, 298)
// For a diagonal Hessian tr(E H E^T) == sum_ij h_j E_ij^2, which is what replaces the
// reference's four dense products per error evaluation.
#include "common.cuh"
#include "internal.h"

namespace cb {

// ---------------------------------------------------------------- vector helpers
template <int VEC> struct FVec;
template <> struct FVec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void load_stream(const float* p) { float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct FVec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void load_stream(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

template <int VEC, typename code_t>
__device__ __forceinline__ void load_codes(const code_t* p, int (&c)[VEC]) {
  if (VEC == 4 && sizeof(code_t) == 1) {
    uint32_t w = *reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int k = 0; k < VEC; ++k) c[k] = (int)(int8_t)((w >> (8 * k)) & 255u);
  } else if (VEC == 4) {
    uint2 w = *reinterpret_cast<const uint2*>(p);
    c[0] = (int)(int16_t)(w.x & 0xFFFFu); c[1 % VEC] = (int)(int16_t)(w.x >> 16);
    c[2 % VEC] = (int)(int16_t)(w.y & 0xFFFFu); c[3 % VEC] = (int)(int16_t)(w.y >> 16);
  } else {
    c[0] = (int)p[0];
  }
}
template <int VEC, typename code_t>
__device__ __forceinline__ void store_codes(code_t* p, const int (&c)[VEC]) {
  if (VEC == 4 && sizeof(code_t) == 1) {
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) w |= (uint32_t)(uint8_t)(int8_t)c[k] << (8 * k);
    *reinterpret_cast<uint32_t*>(p) = w;
  } else if (VEC == 4) {
    uint2 w;
    w.x = (uint32_t)(uint16_t)(int16_t)c[0] | ((uint32_t)(uint16_t)(int16_t)c[1 % VEC] << 16);
    w.y = (uint32_t)(uint16_t)(int16_t)c[2 % VEC] | ((uint32_t)(uint16_t)(int16_t)c[3 % VEC] << 16);
    *reinterpret_cast<uint2*>(p) = w;
  } else {
    p[0] = (code_t)c[0];
  }
}

static inline bool can_vec4(int64_t n, std::initializer_list<const void*> ptrs16) {
  if (n % 4 != 0) return false;
  for (const void* p : ptrs16)
    if (p != nullptr && !aligned16(p)) return false;
  return true;
}

// column index of a chunk with 32-bit arithmetic (the host wrappers require numel < 2^33)
#define CB_CHUNK_COL(VEC, n) ((int64_t)((uint32_t)ch__ % (uint32_t)((n) / (VEC))) * (VEC))
#define CB_GRID_STRIDE_CHUNKS(VEC, numel)                                                 \
  const int64_t nchunk__ = (numel) / (VEC);                                               \
  const int64_t stride__ = (int64_t)gridDim.x * blockDim.x;                               \
  for (int64_t ch__ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch__ < nchunk__; ch__ += stride__)

// ---------------------------------------------------------------- global scale
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t numel, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float part = 0.f;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float v = x[i];
    part = fmaf(v, v, part);
    if (++cnt == 64) { acc += (double)part; part = 0.f; cnt = 0; }
  }
  acc += (double)part;
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void finalize_gs_kernel(const double* sumsq, int64_t numel, float gs_in, int scale_w, float* scalars,
                                   int64_t bstride) {
  sumsq = boff(sumsq, bstride * blockIdx.x); scalars = boff(scalars, bstride * blockIdx.x);
  // alg.py:38-41: W.square().mean().sqrt().item(), fp32
  float gs = 1.f;
  if (scale_w) gs = gs_in > 0.f ? gs_in : sqrtf((float)(sumsq[0] / (double)numel));
  scalars[0] = gs;
  scalars[1] = __int_as_float(0x7f800000);  // min_error = +inf
  scalars[2] = -1.f;                        // best step
  scalars[3] = 0.f; scalars[4] = 0.f; scalars[5] = __int_as_float(0x7f800000); scalars[6] = 0.f; scalars[7] = 0.f;
}

template <int VEC>
__global__ void __launch_bounds__(256)
scale_den_kernel(const float* __restrict__ W, float* __restrict__ Ws, int64_t numel, int64_t n,
                 const float* __restrict__ gs_ptr, const float* __restrict__ h, double* __restrict__ den,
                 float* __restrict__ amax) {
  __shared__ double red[32];
  __shared__ float redf[32];
  const ScaleRecip gs = make_scale_recip(gs_ptr[0]);
  double acc = 0.0;
  float amx = 0.f;      // max |W / gs|: the abs-max of the first Q update (its residual is W / gs itself)
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    const int64_t j = CB_CHUNK_COL(VEC, n);
    FVec<VEC> w, hv;
    w.load_stream(W + i);
    if (h != nullptr) hv.load(h + j);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      w.v[k] = div_by_scale(w.v[k], gs);  // alg.py:42, true (correctly rounded) division
      part = fmaf((h != nullptr ? hv.v[k] : 1.f) * w.v[k], w.v[k], part);
      amx = fmaxf(amx, fabsf(w.v[k]));
    }
    if (Ws != nullptr) w.store(Ws + i);
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(den, acc);
  if (amax != nullptr) {
    amx = block_max(amx, redf);
    if (threadIdx.x == 0) atomic_max_nonneg(amax, amx);
  }
}

// ---------------------------------------------------------------- diagonal Hessian prep
__global__ void __launch_bounds__(1024)
prep_h_kernel(const float* __restrict__ h_in_, int n, float sigma_reg, int aware, float* __restrict__ h_eff_,
              float* __restrict__ sqrt_h_, float* __restrict__ inv_sqrt_h_, float* __restrict__ w_inner_, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.x;
  const float* h_in = boff(h_in_, bo);
  float *h_eff = boff(h_eff_, bo), *sqrt_h = boff(sqrt_h_, bo), *inv_sqrt_h = boff(inv_sqrt_h_, bo), *w_inner = boff(w_inner_, bo);
  __shared__ float red[32];
  __shared__ float s_shift;
  float mn = __int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < n; i += blockDim.x) mn = fminf(mn, h_in != nullptr ? h_in[i] : 1.f);
  // block min via max of negatives
  float neg = -mn;
  neg = warp_max(neg);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = neg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float best = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = fmaxf(best, red[w]);
    const float lo = -best;
    // alg.py:59-63: shift the spectrum up to sigma_reg (aware branch only)
    s_shift = (aware && lo < sigma_reg) ? (sigma_reg - lo) : 0.f;
  }
  __syncthreads();
  const float shift = s_shift;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float h = (h_in != nullptr ? h_in[i] : 1.f) + shift;
    h_eff[i] = h;
    if (aware) {
      const float sh = sqrtf(h);
      sqrt_h[i] = sh;
      inv_sqrt_h[i] = __fdiv_rn(1.f, sh);
      w_inner[i] = h;
    } else {
      sqrt_h[i] = 1.f;
      inv_sqrt_h[i] = 1.f;
      w_inner[i] = h * h;  // H_sqrt := H when not activation aware (alg.py:50)
    }
  }
}

__global__ void __launch_bounds__(256)
hessian_probe_kernel(const float* __restrict__ H, int64_t n, float* __restrict__ diag, int* __restrict__ offdiag_count) {
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int cnt = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / n, j = e - i * n;
    const float v = H[e];
    if (i == j) diag[i] = v;
    else if (v != 0.f) ++cnt;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) atomicAdd(offdiag_count, cnt);
}
__global__ void probe_finish_kernel(int* flag) { flag[0] = (flag[0] == 0) ? 1 : 0; }

// ---------------------------------------------------------------- Q update
template <int VEC>
__global__ void __launch_bounds__(256)
resid_absmax_kernel(const float* __restrict__ Ws, const float* __restrict__ LR, int64_t numel, float* __restrict__ amax) {
  __shared__ float red[32];
  float a = 0.f;
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> w, p;
    w.load(Ws + i);
    if (LR != nullptr) p.load_stream(LR + i);
#pragma unroll
    for (int k = 0; k < VEC; ++k) a = fmaxf(a, fabsf(LR != nullptr ? w.v[k] - p.v[k] : w.v[k]));
  }
  a = block_max(a, red);
  if (threadIdx.x == 0) atomic_max_nonneg(amax, a);
}

template <int VEC, typename code_t>
__global__ void __launch_bounds__(256)
quant_err_kernel(const float* __restrict__ Ws, const float* __restrict__ LR, const float* __restrict__ h,
                 int64_t numel, int64_t n, const float* __restrict__ amax, float eps, float lv,
                 code_t* __restrict__ codes, float* __restrict__ qscale, double* __restrict__ num) {
  __shared__ double red[32];
  const float s = fmaxf(amax[0], eps);
  const ScaleRecip sr = make_scale_recip(s), lvr = make_scale_recip(lv);
  const bool tern = lv == 1.f && ternary_ok(s);      // 2-bit grid: exact codes from two compares (common.cuh)
  const float hs = 0.5f * s;
  if (blockIdx.x == 0 && threadIdx.x == 0) qscale[0] = s;
  double acc = 0.0;
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    const int64_t j = CB_CHUNK_COL(VEC, n);
    FVec<VEC> w, p, hv;
    w.load(Ws + i);
    if (LR != nullptr) p.load_stream(LR + i);
    if (h != nullptr) hv.load(h + j);
    int c[VEC];
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float res = LR != nullptr ? w.v[k] - p.v[k] : w.v[k];
      c[k] = tern ? ternary_code(res, hs) : quant_code(res, sr, lv);
      const float e = res - dequant_val(c[k], s, lvr);
      part = fmaf((h != nullptr ? hv.v[k] : 1.f) * e, e, part);
    }
    store_codes<VEC, code_t>(codes + i, c);
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(num, acc);
}

// ---------------------------------------------------------------- LR update inputs / error
template <int VEC, typename code_t>
__global__ void __launch_bounds__(256)
form_y_kernel(const float* __restrict__ Ws, const code_t* __restrict__ codes, const float* __restrict__ qscale,
              float lv, const float* __restrict__ sqrt_h, int64_t numel, int64_t n,
              float* __restrict__ Y, float* __restrict__ RES) {
  const float s = codes != nullptr ? qscale[0] : 0.f;
  const ScaleRecip lvr = make_scale_recip(lv);
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    const int64_t j = CB_CHUNK_COL(VEC, n);
    FVec<VEC> w, sh, y;
    w.load(Ws + i);
    if (sqrt_h != nullptr) sh.load(sqrt_h + j);
    int c[VEC];
    if (codes != nullptr) load_codes<VEC, code_t>(codes + i, c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (codes != nullptr) w.v[k] = w.v[k] - dequant_val(c[k], s, lvr);
      y.v[k] = sqrt_h != nullptr ? w.v[k] * sh.v[k] : w.v[k];
    }
    if (Y != nullptr) y.store(Y + i);
    if (RES != nullptr) w.store(RES + i);
  }
}

// The codes of one chunk as loaded (unpacked only where they are used, so that several chunks in flight cost one or
// two registers each instead of VEC).
template <int VEC, typename code_t> struct CodeWord;
template <> struct CodeWord<4, int8_t> {
  uint32_t w;
  __device__ __forceinline__ void load(const int8_t* p) { w = *reinterpret_cast<const uint32_t*>(p); }
  __device__ __forceinline__ int get(int k) const { return (int)(int8_t)((w >> (8 * k)) & 255u); }
};
template <> struct CodeWord<4, int16_t> {
  uint2 w;
  __device__ __forceinline__ void load(const int16_t* p) { w = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ int get(int k) const {
    const uint32_t x = k < 2 ? w.x : w.y;
    return (int)(int16_t)((k & 1) ? (x >> 16) : (x & 0xFFFFu));
  }
};
template <typename code_t> struct CodeWord<1, code_t> {
  code_t w;
  __device__ __forceinline__ void load(const code_t* p) { w = p[0]; }
  __device__ __forceinline__ int get(int) const { return (int)w; }
};

// U chunks per thread and iteration, every load of all of them issued before the first use (a lone chunk per
// iteration leaves 36 bytes per thread in flight, which is not enough to cover the HBM latency at full bandwidth);
// four CTAs per SM (<= 64 registers), i.e. ~110 KB of loads in flight per SM.
// UNIT: the ternary grid (levels = 1), where code / levels * scale is code * scale and no division has to be emulated.
template <int VEC, typename code_t, bool UNIT>
__global__ void __launch_bounds__(256, 4)
err_kernel(const float* __restrict__ Ws, const code_t* __restrict__ codes, const float* __restrict__ qscale,
           float lv, const float* __restrict__ LR, const float* __restrict__ wcol, int64_t numel, int64_t n,
           double* __restrict__ num, float* __restrict__ amax_next) {
  constexpr int U = VEC == 4 ? 3 : 1;
  __shared__ double red[32];
  __shared__ float redf[32];
  const float s = codes != nullptr ? qscale[0] : 0.f;
  const ScaleRecip lvr = make_scale_recip(lv);
  double acc = 0.0;
  float amx = 0.f;   // max |Ws - LR|: the abs-max the next Q update would otherwise need a pass of its own for
  const int64_t nchunk = numel / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t chunks_per_row = (uint32_t)(n / VEC);
  for (int64_t ch0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch0 < nchunk; ch0 += U * stride) {
    FVec<VEC> w[U], p[U];
    CodeWord<VEC, code_t> cw[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t ch = ch0 + u * stride;
      ok[u] = ch < nchunk;
      if (ok[u]) {
        const int64_t i = ch * VEC;
        w[u].load(Ws + i);
        if (LR != nullptr) p[u].load_stream(LR + i);
        if (codes != nullptr) cw[u].load(codes + i);
      }
    }
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ok[u]) {
        FVec<VEC> hv;        // column weights: a 4 n byte vector that lives in L1, fetched where it is used
        if (wcol != nullptr) hv.load(wcol + (int64_t)((uint32_t)(ch0 + u * stride) % chunks_per_row) * VEC);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float e = w[u].v[k];
          if (codes != nullptr) e -= UNIT ? __fmul_rn((float)cw[u].get(k), s) : dequant_val(cw[u].get(k), s, lvr);
          if (LR != nullptr) e -= p[u].v[k];
          part = fmaf((wcol != nullptr ? hv.v[k] : 1.f) * e, e, part);
          if (amax_next != nullptr) amx = fmaxf(amx, fabsf(LR != nullptr ? w[u].v[k] - p[u].v[k] : w[u].v[k]));
        }
      }
    }
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(num, acc);
  if (amax_next != nullptr) {
    amx = block_max(amx, redf);
    if (threadIdx.x == 0) atomic_max_nonneg(amax_next, amx);
  }
}

// ---------------------------------------------------------------- fused residual -> bf16 operands
// One pass for the tensor-core rank-r step: res = Ws - dequant(codes), Y = res (.) sqrt(h) written as
// the two K-major bf16 operands Yb (m x n) and Ytb (n x m, transposed through a shared-memory tile),
// plus the fp32 residual when the LPLR loop needs it.  Replaces form_y + two conversion passes.
template <typename code_t>
__global__ void __launch_bounds__(256)
form_y_bf16_kernel(const float* __restrict__ Ws, const code_t* __restrict__ codes, const float* __restrict__ qscale,
                   float lv, const float* __restrict__ sqrt_h, int m, int n, __nv_bfloat16* __restrict__ Yb,
                   __nv_bfloat16* __restrict__ Ytb, float* __restrict__ RES) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][68];   // [col][row]
  const float s = codes != nullptr ? qscale[0] : 0.f;
  const ScaleRecip lvr = make_scale_recip(lv);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int col = blockIdx.x * 64 + 4 * tx;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rl = ty + 16 * k, row = blockIdx.y * 64 + rl;
    float y[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < m && col < n) {
      const int64_t i = (int64_t)row * n + col;
      FVec<4> w, sh;
      w.load_stream(Ws + i);
      if (sqrt_h != nullptr) sh.load(sqrt_h + col);
      int c[4];
      if (codes != nullptr) load_codes<4, code_t>(codes + i, c);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (codes != nullptr) w.v[j] -= dequant_val(c[j], s, lvr);
        y[j] = sqrt_h != nullptr ? w.v[j] * sh.v[j] : w.v[j];
      }
      if (RES != nullptr) w.store(RES + i);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]), p1 = __floats2bfloat162_rn(y[2], y[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(Yb + i) = pk;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[4 * tx + j][rl] = __float2bfloat16_rn(y[j]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cl = ty + 16 * k;                 // column of the tile = row of Ytb
    const int gc = blockIdx.x * 64 + cl, gr = blockIdx.y * 64 + 4 * tx;
    if (gc < n && gr < m) {
      const uint2 pk = *reinterpret_cast<const uint2*>(&tile[cl][4 * tx]);
      *reinterpret_cast<uint2*>(Ytb + (int64_t)gc * m + gr) = pk;
    }
  }
}

// Q update and operand builder in one pass (tensor-core path): quantise res = Ws - LR with the scale the
// abs-max pass delivered, store the codes, accumulate the weighted error of the new iterate, and write
// Y = (Ws - Q) (.) sqrt(h) as both bf16 operands (+ the fp32 residual for the LPLR loop).  Same element
// arithmetic as quant_err_kernel followed by form_y_bf16_kernel, one read of Ws instead of two and no
// read-back of the codes.
// 1.0f / 0.0f comparison results (FSET.BF)
__device__ __forceinline__ float fset_gt_f(float a, float b) {
  float d;
  asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fset_lt_f(float a, float b) {
  float d;
  asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// The 64 x 64 bf16 tile that turns Y into its transpose lives in shared memory as tile[col][row] with rows of 64
// elements (16 units of 8 bytes) and the unit index XOR-ed with (col / 4) % 16: a thread writes one unit (4 consecutive
// rows of one column) per column it owns and reads one unit per row of Ytb it writes, and both patterns touch every
// bank pair exactly twice per warp instruction (the minimum for 256 bytes).
__device__ __forceinline__ int tile_unit(int col, int unit) { return col * 16 + (unit ^ ((col >> 2) & 15)); }

// TERN: the ternary grid with a scale inside the validity range of the compare-only code (common.cuh).  The code is
// cf = (res > s/2) - (res < -s/2) as a float, so its dequantised value is the exact product cf * s (levels = 1: the
// division by the level count is the identity) and neither the quantiser's nor the dequantiser's division is emulated.
template <typename code_t, bool TERN>
__device__ __forceinline__ void quant_form_y_body(const float* __restrict__ Ws, const float* __restrict__ LR,
                                                  const float* __restrict__ h_err, const float* __restrict__ sqrt_h,
                                                  const int m, const int n, const float s, const float lv,
                                                  code_t* __restrict__ codes, double* __restrict__ num,
                                                  __nv_bfloat16* __restrict__ Yb, __nv_bfloat16* __restrict__ Ytb,
                                                  float* __restrict__ RES, uint2* tile, double* red) {
  const ScaleRecip sr = make_scale_recip(s), lvr = make_scale_recip(lv);
  const float hs = 0.5f * s, nhs = -hs;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int col = blockIdx.x * 64 + 4 * tx;
  const bool col_ok = col < n;
  FVec<4> he, sh;
  if (h_err != nullptr && col_ok) he.load(h_err + col);
  if (sqrt_h != nullptr && col_ok) sh.load(sqrt_h + col);
  double acc = 0.0;
  float y[4][4];                 // [row k of the thread's 4 consecutive rows][column j]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = blockIdx.y * 64 + 4 * ty + k;
#pragma unroll
    for (int j = 0; j < 4; ++j) y[k][j] = 0.f;
    if (row < m && col_ok) {
      const int64_t i = (int64_t)row * n + col;
      FVec<4> w, p;
      w.load_stream(Ws + i);
      if (LR != nullptr) p.load_stream(LR + i);
      int c[4];
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float res = LR != nullptr ? w.v[j] - p.v[j] : w.v[j];
        float dq;
        if (TERN) {
          const float cf = fset_gt_f(res, hs) - fset_lt_f(res, nhs);
          c[j] = __float2int_rn(cf);
          dq = __fmul_rn(cf, s);
        } else {
          c[j] = quant_code(res, sr, lv);
          dq = dequant_val(c[j], s, lvr);
        }
        const float e = res - dq;
        part = fmaf((h_err != nullptr ? he.v[j] : 1.f) * e, e, part);
        w.v[j] -= dq;                                      // Ws - Q
        y[k][j] = sqrt_h != nullptr ? w.v[j] * sh.v[j] : w.v[j];
      }
      acc += (double)part;
      store_codes<4, code_t>(codes + i, c);
      if (RES != nullptr) w.store(RES + i);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(y[k][0], y[k][1]), p1 = __floats2bfloat162_rn(y[k][2], y[k][3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(Yb + i) = pk;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {          // column 4 tx + j, rows 4 ty .. 4 ty + 3: one 8-byte unit
    __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0][j], y[1][j]), p1 = __floats2bfloat162_rn(y[2][j], y[3][j]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    tile[tile_unit(4 * tx + j, ty)] = pk;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cl = ty + 16 * k;                 // column of the tile = row of Ytb
    const int gc = blockIdx.x * 64 + cl, gr = blockIdx.y * 64 + 4 * tx;
    if (gc < n && gr < m) *reinterpret_cast<uint2*>(Ytb + (int64_t)gc * m + gr) = tile[tile_unit(cl, tx)];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(num, acc);
}

template <typename code_t>
__global__ void __launch_bounds__(256)
quant_form_y_bf16_kernel(const float* __restrict__ Ws, const float* __restrict__ LR, const float* __restrict__ h_err,
                         const float* __restrict__ sqrt_h, int m, int n, const float* __restrict__ amax, float eps,
                         float lv, code_t* __restrict__ codes, float* __restrict__ qscale, double* __restrict__ num,
                         __nv_bfloat16* __restrict__ Yb, __nv_bfloat16* __restrict__ Ytb, float* __restrict__ RES) {
  __shared__ __align__(16) uint2 tile[64 * 16];          // [col][unit of 4 rows], swizzled (tile_unit)
  __shared__ double red[32];
  const float s = fmaxf(amax[0], eps);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) qscale[0] = s;
  if (lv == 1.f && ternary_ok(s))      // 2-bit grid: exact codes from two compares (common.cuh)
    quant_form_y_body<code_t, true>(Ws, LR, h_err, sqrt_h, m, n, s, lv, codes, num, Yb, Ytb, RES, tile, red);
  else
    quant_form_y_body<code_t, false>(Ws, LR, h_err, sqrt_h, m, n, s, lv, codes, num, Yb, Ytb, RES, tile, red);
}

// ---------------------------------------------------------------- dense-Hessian helpers
// E = Ws - Q - L R written out (the dense metric tr(E H E^T) needs E as a GEMM operand)
template <int VEC, typename code_t>
__global__ void __launch_bounds__(256)
form_e_kernel(const float* __restrict__ Ws, const code_t* __restrict__ codes, const float* __restrict__ qscale,
              float lv, const float* __restrict__ LR, int64_t numel, float* __restrict__ E) {
  const float s = codes != nullptr ? qscale[0] : 0.f;
  const ScaleRecip lvr = make_scale_recip(lv);
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> w, p;
    w.load(Ws + i);
    if (LR != nullptr) p.load_stream(LR + i);
    int c[VEC];
    if (codes != nullptr) load_codes<VEC, code_t>(codes + i, c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (codes != nullptr) w.v[k] -= dequant_val(c[k], s, lvr);
      if (LR != nullptr) w.v[k] -= p.v[k];
    }
    w.store(E + i);
  }
}

// out += sum_i a_i * b_i  (b == a gives a squared Frobenius norm)
__global__ void __launch_bounds__(256)
dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t numel, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  float part = 0.f;
  int cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    part = fmaf(a[i], b[i], part);
    if (++cnt == 32) { acc += (double)part; part = 0.f; cnt = 0; }
  }
  acc += (double)part;
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// Hs = (H + H^T) / 2  (alg.py:54)
__global__ void __launch_bounds__(256)
symmetrize_kernel(const float* __restrict__ H, int64_t n, float* __restrict__ Hs) {
  const int64_t total = n * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / n, j = e - i * n;
    Hs[e] = (H[e] + H[j * n + i]) / 2.f;
  }
}

// ---------------------------------------------------------------- Convex-CALDERA prox steps
// Extrapolation point, gradient of the smooth term and the two prox arguments in one pass
// (diagonal H):  Y = X + beta (X - Xprev);  G = (W - Y_L - Y_R) (.) h;  V = Y + t G.
// Also accumulates ||V_R||_F^2 for the radial shrink of the R block.
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_point_kernel(const float* __restrict__ W, const float* __restrict__ L, const float* __restrict__ Lp,
                 const float* __restrict__ R, const float* __restrict__ Rp, const float* __restrict__ h,
                 int64_t numel, int64_t n, float beta, float t, float* __restrict__ VL, float* __restrict__ VR,
                 double* __restrict__ vr_sumsq) {
  __shared__ double red[32];
  double acc = 0.0;
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    const int64_t j = CB_CHUNK_COL(VEC, n);
    FVec<VEC> w, l, lp, r, rp, hv, vl, vr;
    w.load_stream(W + i); l.load(L + i); lp.load_stream(Lp + i); r.load(R + i); rp.load_stream(Rp + i);
    if (h != nullptr) hv.load(h + j);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float yl = l.v[k] + beta * (l.v[k] - lp.v[k]);
      const float yr = r.v[k] + beta * (r.v[k] - rp.v[k]);
      const float g = (w.v[k] - yl - yr) * (h != nullptr ? hv.v[k] : 1.f);
      vl.v[k] = fmaf(t, g, yl);
      vr.v[k] = fmaf(t, g, yr);
      part = fmaf(vr.v[k], vr.v[k], part);
    }
    vl.store(VL + i);
    vr.store(VR + i);
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(vr_sumsq, acc);
}

// sc[]: [0] nuclear norm of the new L, [1] alpha, [2] ||R_new||^2; sigma2[] are squared singular
// values of V_L (descending).  Writes the column weights w_i = shrink(sigma_i) / sqrt(sigma_i) ... see below.
// The factors handed over are Lf = U sqrt(S), Rf = sqrt(S) V^T, so  U diag(s') V^T = (Lf diag(s'/S)) Rf.
__global__ void __launch_bounds__(512)
cvx_shrink_kernel(const float* __restrict__ sigma2, int r, float thresh, float tau_star, int constrained,
                  const double* __restrict__ vr_sumsq, float t_lambda, float kappa, float q0,
                  float* __restrict__ colw, float* __restrict__ s_out, double* __restrict__ sc) {
  __shared__ float s_sig[512];
  __shared__ float s_theta;
  const int tid = threadIdx.x;
  if (tid < r) s_sig[tid] = sqrtf(fmaxf(sigma2[tid], 0.f));
  __syncthreads();
  if (tid == 0) {
    float theta = thresh;
    if (constrained) {
      // projection of the (descending, non-negative) singular values onto {sum <= tau_star}
      double css = 0.0;
      for (int i = 0; i < r; ++i) css += s_sig[i];
      theta = 0.f;
      if (css > (double)tau_star) {
        double run = 0.0;
        int kk = 0;
        for (int i = 0; i < r; ++i) {
          run += s_sig[i];
          if ((double)s_sig[i] * (i + 1) > run - (double)tau_star) kk = i;
        }
        run = 0.0;
        for (int i = 0; i <= kk; ++i) run += s_sig[i];
        theta = (float)((run - (double)tau_star) / (double)(kk + 1));
      }
    }
    s_theta = theta;
  }
  __syncthreads();
  float mine = 0.f;
  if (tid < r) {
    const float sg = s_sig[tid];
    mine = fmaxf(sg - s_theta, 0.f);
    colw[tid] = sg > 0.f ? mine / sg : 0.f;
    s_out[tid] = mine;
  }
  __shared__ float red[16];
  float tot = warp_sum(mine);
  if ((tid & 31) == 0) red[tid >> 5] = tot;
  __syncthreads();
  if (tid == 0) {
    double nuc = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) nuc += red[w];
    // radial prox of t*lambda*max(q0, ||R||^2 / kappa) at V_R
    const double v = vr_sumsq[0];
    double alpha = 1.0;
    if (v / kappa > q0) {
      const double a1 = 1.0 / (1.0 + 2.0 * t_lambda / kappa);
      alpha = (a1 * a1 * v / kappa >= q0) ? a1 : sqrt((double)q0 * kappa / v);
    }
    sc[0] = nuc;
    sc[1] = alpha;
    sc[2] = alpha * alpha * v;
  }
}

// R_new = alpha V_R, and the smooth term 1/2 sum_j h_j (W - L_new - R_new)^2
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_finish_kernel(const float* __restrict__ W, const float* __restrict__ Lnew, const float* __restrict__ VR,
                  const float* __restrict__ h, int64_t numel, int64_t n, const double* __restrict__ sc,
                  float* __restrict__ Rnew, double* __restrict__ smooth) {
  __shared__ double red[32];
  const float alpha = (float)sc[1];
  double acc = 0.0;
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    const int64_t j = CB_CHUNK_COL(VEC, n);
    FVec<VEC> w, l, v, hv;
    w.load_stream(W + i); l.load(Lnew + i); v.load_stream(VR + i);
    if (h != nullptr) hv.load(h + j);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      v.v[k] *= alpha;
      const float e = w.v[k] - l.v[k] - v.v[k];
      part = fmaf((h != nullptr ? hv.v[k] : 1.f) * e, e, part);
    }
    v.store(Rnew + i);
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(smooth, 0.5 * acc);
}

// Dense Hessian: the gradient (W - Y_L - Y_R) H is a contraction, so the step is split around it.
// Part 1: D = W - Y_L - Y_R at the extrapolated point Y = X + beta (X - X_prev).
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_resid_kernel(const float* __restrict__ W, const float* __restrict__ L, const float* __restrict__ Lp,
                 const float* __restrict__ R, const float* __restrict__ Rp, int64_t numel, float beta,
                 float* __restrict__ D) {
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> w, l, lp, r, rp, d;
    w.load_stream(W + i); l.load(L + i); lp.load(Lp + i); r.load(R + i); rp.load(Rp + i);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float yl = l.v[k] + beta * (l.v[k] - lp.v[k]);
      const float yr = r.v[k] + beta * (r.v[k] - rp.v[k]);
      d.v[k] = w.v[k] - yl - yr;
    }
    d.store(D + i);
  }
}
// Part 2: VL holds G = D H on entry; V_L = Y_L + t G (in place), V_R = Y_R + t G, sum V_R^2.
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_point_dense_kernel(const float* __restrict__ L, const float* __restrict__ Lp, const float* __restrict__ R,
                       const float* __restrict__ Rp, int64_t numel, float beta, float t, float* __restrict__ VL,
                       float* __restrict__ VR, double* __restrict__ vr_sumsq) {
  __shared__ double red[32];
  double acc = 0.0;
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> l, lp, r, rp, g, vl, vr;
    l.load(L + i); lp.load_stream(Lp + i); r.load(R + i); rp.load_stream(Rp + i); g.load(VL + i);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float yl = l.v[k] + beta * (l.v[k] - lp.v[k]);
      const float yr = r.v[k] + beta * (r.v[k] - rp.v[k]);
      vl.v[k] = fmaf(t, g.v[k], yl);
      vr.v[k] = fmaf(t, g.v[k], yr);
      part = fmaf(vr.v[k], vr.v[k], part);
    }
    vl.store(VL + i);
    vr.store(VR + i);
    acc += (double)part;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(vr_sumsq, acc);
}
// R_new = alpha V_R and (optionally) E = W - L_new - R_new for the smooth term 1/2 sum (E H) (.) E
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_finish_dense_kernel(const float* __restrict__ W, const float* __restrict__ Lnew, const float* __restrict__ VR,
                        int64_t numel, const double* __restrict__ sc, float* __restrict__ Rnew, float* __restrict__ E) {
  const float alpha = (float)sc[1];
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> w, l, v, e;
    v.load_stream(VR + i);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v.v[k] *= alpha;
    v.store(Rnew + i);
    if (E != nullptr) {
      w.load_stream(W + i); l.load(Lnew + i);
#pragma unroll
      for (int k = 0; k < VEC; ++k) e.v[k] = w.v[k] - l.v[k] - v.v[k];
      e.store(E + i);
    }
  }
}
__global__ void scale_double_kernel(double* x, double f) { x[0] *= f; }

// quantize_residual (convex_caldera.py:342-373): delta = 2 t / (2^b - 1) (t / 2^15 for b = 16),
// R_int = clamp(rint(R / delta), +-(2^(b-1) - 1)); out = base + delta * R_int
template <int VEC>
__global__ void __launch_bounds__(256)
cvx_quant_residual_kernel(const float* __restrict__ R, const float* __restrict__ base, int64_t numel, const float* __restrict__ amax,
                          int bits, float* __restrict__ Rq, float* __restrict__ Wc, float* __restrict__ delta_out) {
  const float tmax = amax[0];
  const float delta = bits < 16 ? __fdiv_rn(2.f * tmax, (float)((1 << bits) - 1)) : __fdiv_rn(tmax, 32768.f);
  const float mx = (float)((1 << (bits - 1)) - 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) delta_out[0] = delta;
  const ScaleRecip dr = make_scale_recip(delta);
  CB_GRID_STRIDE_CHUNKS(VEC, numel) {
    const int64_t i = ch__ * VEC;
    FVec<VEC> r, b, q, wc;
    r.load_stream(R + i);
    if (base != nullptr) b.load_stream(base + i);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float c = rintf(div_by_scale(r.v[k], dr));
      c = fminf(fmaxf(c, -mx), mx);
      q.v[k] = __fmul_rn(delta, c);
      wc.v[k] = base != nullptr ? b.v[k] + q.v[k] : q.v[k];
    }
    q.store(Rq + i);
    if (Wc != nullptr) wc.store(Wc + i);
  }
}

// out[0] += sum x, out[1] += sum x^2, out[2] += sum (x - y)^2 (y may be null)
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t numel, double* __restrict__ out) {
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const double v = (double)x[i];
    s1 += v;
    s2 += v * v;
    if (y != nullptr) { const double d = v - (double)y[i]; s3 += d * d; }
  }
  s1 = block_sum(s1, red); s2 = block_sum(s2, red); s3 = block_sum(s3, red);
  if (threadIdx.x == 0) { atomicAdd(out, s1); atomicAdd(out + 1, s2); atomicAdd(out + 2, s3); }
}

// ---------------------------------------------------------------- selection / bookkeeping
__global__ void select_outer_kernel(double* num, const double* den, float* errors, int step, float* scalars,
                                    int* flags, int all_updated, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.x;
  num = boff(num, bo); den = boff(den, bo); errors = boff(errors, bo); scalars = boff(scalars, bo); flags = boff(flags, bo);
  // alg.py:104-107: strict '<' and only once every component has been updated
  const float err = (float)sqrt(num[0] / den[0]);
  errors[step] = err;
  const int take = (err < scalars[1]) && all_updated;
  flags[0] = take;
  if (take) { scalars[1] = err; scalars[2] = (float)step; }
  num[0] = 0.0;
}
__global__ void select_inner_kernel(double* num, float* scalars, int* flags, int first, int last, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.x;
  num = boff(num, bo); scalars = boff(scalars, bo); flags = boff(flags, bo);
  // alg.py:182-188
  const float err = (float)sqrt(num[0]);
  const float best = first ? __int_as_float(0x7f800000) : scalars[5];
  const int take = err < best;
  flags[1] = take;
  if (take) scalars[5] = err; else if (first) scalars[5] = best;
  // An LR update in which no inner iterate was ever taken (every error NaN) leaves best_L_quant_out = None in
  // the reference, which then raises (alg.py:190); count those updates so the caller can raise too.
  const int any = (first ? 0 : flags[6]) | take;
  flags[6] = any;
  if (last && !any) flags[5] += 1;
  num[0] = 0.0;
}

__global__ void __launch_bounds__(256)
copy_if_kernel(const int* __restrict__ flag, uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes, int vec_ok) {
  if (flag != nullptr && flag[0] == 0) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const size_t n16 = bytes >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t i = t0; i < n16; i += stride) d4[i] = s4[i];
    for (size_t i = (n16 << 4) + t0; i < bytes; i += stride) dst[i] = src[i];
  } else {
    for (size_t i = t0; i < bytes; i += stride) dst[i] = src[i];
  }
}

// best_decomp = deepcopy(curr_decomp) for all its parts in one launch: blockIdx.y selects the segment
__global__ void __launch_bounds__(256) copy_if_multi_kernel(const int* __restrict__ flag, const CopySegments segs, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.z;
  if (flag != nullptr && boff(flag, bo)[0] == 0) return;
  const int k = blockIdx.y;
  const size_t bytes = segs.bytes[k];
  uint8_t* __restrict__ dst = boff(reinterpret_cast<uint8_t*>(segs.dst[k]), bo);
  const uint8_t* __restrict__ src = boff(reinterpret_cast<const uint8_t*>(segs.src[k]), bo);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (src == nullptr) {                      // zero fill
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      const size_t n16 = bytes >> 4;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (size_t i = t0; i < n16; i += stride) d4[i] = make_uint4(0, 0, 0, 0);
      for (size_t i = (n16 << 4) + t0; i < bytes; i += stride) dst[i] = 0;
    } else {
      for (size_t i = t0; i < bytes; i += stride) dst[i] = 0;
    }
    return;
  }
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
    const size_t n16 = bytes >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t i = t0; i < n16; i += stride) d4[i] = s4[i];
    for (size_t i = (n16 << 4) + t0; i < bytes; i += stride) dst[i] = src[i];
  } else {
    for (size_t i = t0; i < bytes; i += stride) dst[i] = src[i];
  }
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ p, int64_t count, uint64_t seed,
                                                    const uint64_t* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed += seed_dev[0];   // per-layer seed kept in device memory (CUDA-graph replays)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    p[i] = gaussian_from(seed, (uint64_t)i);
}

// mode 0: multiply by v, 1: divide by v, 2: multiply by sqrt(sqrt(v)), 3: divide by sqrt(sqrt(v))
// (modes 2/3 take squared singular values, i.e. eigenvalues of the Gram matrix)
__device__ __forceinline__ float apply_mode(float x, float v, int mode) {
  switch (mode) {
    case 0: return x * v;
    case 1: return __fdiv_rn(x, v);
    case 2: return x * sqrtf(sqrtf(fmaxf(v, 0.f)));
    default: { const float d = sqrtf(sqrtf(fmaxf(v, 0.f))); return d > 0.f ? __fdiv_rn(x, d) : 0.f; }
  }
}
__global__ void __launch_bounds__(256)
scale_cols_kernel(const float* __restrict__ X, int64_t rows, int64_t cols, const float* __restrict__ v, int mode, float* __restrict__ out) {
  const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) out[i] = apply_mode(X[i], v[i % cols], mode);
}
__global__ void __launch_bounds__(256)
scale_rows_kernel(const float* __restrict__ X, int64_t rows, int64_t cols, const float* __restrict__ v, int mode, float* __restrict__ out) {
  const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) out[i] = apply_mode(X[i], v[i / cols], mode);
}

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, int64_t rows, int64_t cols, T* __restrict__ dst) {
  __shared__ T tile[32][33];
  const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = by + k, c = bx + tx;
    if (r < rows && c < cols) tile[k][tx] = src[r * cols + c];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t c = bx + k, r = by + tx;  // dst is cols x rows
    if (r < rows && c < cols) dst[c * rows + r] = tile[tx][k];
  }
}

// ---------------------------------------------------------------- host wrappers
int sumsq(const float* x, int64_t numel, double* out, cudaStream_t st) {
  sumsq_kernel<<<grid_for(numel, 256 * 8, 4), 256, 0, st>>>(x, numel, out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int finalize_global_scale(const double* ss, int64_t numel, float gs_in, int scale_w, float* scalars, cudaStream_t st, const Bt& bt) {
  finalize_gs_kernel<<<bt.n, 1, 0, st>>>(ss, numel, gs_in, scale_w, scalars, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int scale_and_den(const float* W, float* Ws, int64_t m, int64_t n, const float* gs, const float* h, double* den, cudaStream_t st,
                  float* amax) {
  const int64_t numel = m * n;
  if (can_vec4(n, {W, Ws, h}))
    scale_den_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(W, Ws, numel, n, gs, h, den, amax);
  else
    scale_den_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(W, Ws, numel, n, gs, h, den, amax);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int prep_hessian_diag(const float* h_in, int64_t n, float sigma_reg, int aware, float* h_eff, float* sqrt_h,
                      float* inv_sqrt_h, float* w_inner, float* scratch, cudaStream_t st, const Bt& bt) {
  (void)scratch;
  prep_h_kernel<<<bt.n, 1024, 0, st>>>(h_in, (int)n, sigma_reg, aware, h_eff, sqrt_h, inv_sqrt_h, w_inner, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int resid_absmax(const float* Ws, const float* LR, int64_t numel, float* amax, cudaStream_t st) {
  CB_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), st));
  if (can_vec4(numel, {Ws, LR}))
    resid_absmax_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(Ws, LR, numel, amax);
  else
    resid_absmax_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(Ws, LR, numel, amax);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int quant_err(const float* Ws, const float* LR, const float* h, int64_t m, int64_t n, const float* amax, float eps,
              int bits, void* codes, float* qscale, double* num, cudaStream_t st) {
  const int64_t numel = m * n;
  const float lv = (float)((1 << (bits - 1)) - 1);
  const bool v4 = can_vec4(n, {Ws, LR, h, codes});
#define CB_QE(VEC, T, G)                                                                                       \
  quant_err_kernel<VEC, T><<<G, 256, 0, st>>>(Ws, LR, h, numel, n, amax, eps, lv, reinterpret_cast<T*>(codes), \
                                              qscale, num)
  if (bits <= 8) { if (v4) CB_QE(4, int8_t, grid_for(numel / 4, 256 * 2, 8)); else CB_QE(1, int8_t, grid_for(numel, 256 * 4, 8)); }
  else { if (v4) CB_QE(4, int16_t, grid_for(numel / 4, 256 * 2, 8)); else CB_QE(1, int16_t, grid_for(numel, 256 * 4, 8)); }
#undef CB_QE
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int form_y(const float* Ws, const void* codes, int bits, const float* qscale, const float* sqrt_h, int64_t m,
           int64_t n, float* Y, float* RES, cudaStream_t st) {
  const int64_t numel = m * n;
  const float lv = (float)((1 << (bits - 1)) - 1);
  const bool v4 = can_vec4(n, {Ws, codes, sqrt_h, Y, RES});
#define CB_FY(VEC, T, G) \
  form_y_kernel<VEC, T><<<G, 256, 0, st>>>(Ws, reinterpret_cast<const T*>(codes), qscale, lv, sqrt_h, numel, n, Y, RES)
  if (bits <= 8) { if (v4) CB_FY(4, int8_t, grid_for(numel / 4, 256 * 2, 8)); else CB_FY(1, int8_t, grid_for(numel, 256 * 4, 8)); }
  else { if (v4) CB_FY(4, int16_t, grid_for(numel / 4, 256 * 2, 8)); else CB_FY(1, int16_t, grid_for(numel, 256 * 4, 8)); }
#undef CB_FY
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int err_accum(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR, const float* w,
              int64_t m, int64_t n, double* num, cudaStream_t st, float* amax_next) {
  const int64_t numel = m * n;
  if (amax_next != nullptr) CB_CUDA(cudaMemsetAsync(amax_next, 0, sizeof(float), st));
  const float lv = (float)((1 << (bits - 1)) - 1);
  const bool v4 = can_vec4(n, {Ws, codes, LR, w});
#define CB_ER(VEC, T, G)                                                                                                          \
  do {                                                                                                                            \
    if (lv == 1.f) err_kernel<VEC, T, true><<<G, 256, 0, st>>>(Ws, reinterpret_cast<const T*>(codes), qscale, lv, LR, w, numel, n, num, amax_next);  \
    else err_kernel<VEC, T, false><<<G, 256, 0, st>>>(Ws, reinterpret_cast<const T*>(codes), qscale, lv, LR, w, numel, n, num, amax_next);           \
  } while (0)
  if (bits <= 8) { if (v4) CB_ER(4, int8_t, grid_for(numel / 4, 256 * 3, 4)); else CB_ER(1, int8_t, grid_for(numel, 256 * 4, 4)); }
  else { if (v4) CB_ER(4, int16_t, grid_for(numel / 4, 256 * 3, 4)); else CB_ER(1, int16_t, grid_for(numel, 256 * 4, 4)); }
#undef CB_ER
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int select_outer(double* num, const double* den, float* errors, int step, float* scalars, int* flags, int all_updated, cudaStream_t st,
                 const Bt& bt) {
  select_outer_kernel<<<bt.n, 1, 0, st>>>(num, den, errors, step, scalars, flags, all_updated, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int select_inner(double* num, float* scalars, int* flags, int first, int last, cudaStream_t st, const Bt& bt) {
  select_inner_kernel<<<bt.n, 1, 0, st>>>(num, scalars, flags, first, last, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int copy_if(const int* flag, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return CB_OK;
  const int vec_ok = aligned16(dst) && aligned16(src);
  copy_if_kernel<<<grid_for((int64_t)(bytes / 16 + 1), 256 * 4, 4), 256, 0, st>>>(
      flag, reinterpret_cast<uint8_t*>(dst), reinterpret_cast<const uint8_t*>(src), bytes, vec_ok);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int copy_if_multi(const int* flag, const CopySegments& segs, cudaStream_t st, const Bt& bt) {
  if (segs.count <= 0) return CB_OK;
  size_t biggest = 0;
  for (int k = 0; k < segs.count; ++k) biggest = segs.bytes[k] > biggest ? segs.bytes[k] : biggest;
  if (biggest == 0) return CB_OK;
  dim3 grid((unsigned)grid_for((int64_t)(biggest / 16 + 1), 256 * 4, bt.n > 1 ? 1 : 4), (unsigned)segs.count, (unsigned)bt.n);
  copy_if_multi_kernel<<<grid, 256, 0, st>>>(flag, segs, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int fill_randn(float* p, int64_t count, uint64_t seed, const uint64_t* seed_dev, cudaStream_t st) {
  randn_kernel<<<grid_for(count, 256 * 4, 4), 256, 0, st>>>(p, count, seed, seed_dev);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int scale_cols(const float* X, int64_t rows, int64_t cols, const float* v, int mode, float* out, cudaStream_t st) {
  scale_cols_kernel<<<grid_for(rows * cols, 256 * 4, 4), 256, 0, st>>>(X, rows, cols, v, mode, out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int scale_rows(const float* X, int64_t rows, int64_t cols, const float* v, int mode, float* out, cudaStream_t st) {
  scale_rows_kernel<<<grid_for(rows * cols, 256 * 4, 4), 256, 0, st>>>(X, rows, cols, v, mode, out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int transpose_codes(const void* src, int64_t rows, int64_t cols, int elem_bytes, void* dst, cudaStream_t st) {
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (elem_bytes == 1) transpose_kernel<int8_t><<<grid, 256, 0, st>>>((const int8_t*)src, rows, cols, (int8_t*)dst);
  else if (elem_bytes == 2) transpose_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)src, rows, cols, (int16_t*)dst);
  else transpose_kernel<float><<<grid, 256, 0, st>>>((const float*)src, rows, cols, (float*)dst);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

int form_e(const float* Ws, const void* codes, int bits, const float* qscale, const float* LR, int64_t m, int64_t n,
           float* E, cudaStream_t st) {
  const int64_t numel = m * n;
  const float lv = (float)((1 << (bits - 1)) - 1);
  const bool v4 = can_vec4(n, {Ws, codes, LR, E});
#define CB_FE(VEC, T, G) \
  form_e_kernel<VEC, T><<<G, 256, 0, st>>>(Ws, reinterpret_cast<const T*>(codes), qscale, lv, LR, numel, E)
  if (bits <= 8) { if (v4) CB_FE(4, int8_t, grid_for(numel / 4, 256 * 2, 8)); else CB_FE(1, int8_t, grid_for(numel, 256 * 4, 8)); }
  else { if (v4) CB_FE(4, int16_t, grid_for(numel / 4, 256 * 2, 8)); else CB_FE(1, int16_t, grid_for(numel, 256 * 4, 8)); }
#undef CB_FE
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int dot_accum(const float* a, const float* b, int64_t numel, double* out, cudaStream_t st) {
  dot_kernel<<<grid_for(numel, 256 * 8, 4), 256, 0, st>>>(a, b, numel, out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int symmetrize(const float* H, int64_t n, float* Hs, cudaStream_t st) {
  symmetrize_kernel<<<grid_for(n * n, 256 * 4, 4), 256, 0, st>>>(H, n, Hs);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

int cvx_point(const float* W, const float* L, const float* Lp, const float* R, const float* Rp, const float* h,
              int64_t m, int64_t n, float beta, float t, float* VL, float* VR, double* vr_sumsq, cudaStream_t st) {
  const int64_t numel = m * n;
  if (can_vec4(n, {W, L, Lp, R, Rp, h, VL, VR}))
    cvx_point_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(W, L, Lp, R, Rp, h, numel, n, beta, t, VL, VR, vr_sumsq);
  else
    cvx_point_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(W, L, Lp, R, Rp, h, numel, n, beta, t, VL, VR, vr_sumsq);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int cvx_resid(const float* W, const float* L, const float* Lp, const float* R, const float* Rp, int64_t m, int64_t n,
              float beta, float* D, cudaStream_t st) {
  const int64_t numel = m * n;
  if (can_vec4(n, {W, L, Lp, R, Rp, D}))
    cvx_resid_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(W, L, Lp, R, Rp, numel, beta, D);
  else
    cvx_resid_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(W, L, Lp, R, Rp, numel, beta, D);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int cvx_point_dense(const float* L, const float* Lp, const float* R, const float* Rp, int64_t m, int64_t n, float beta,
                    float t, float* VL, float* VR, double* vr_sumsq, cudaStream_t st) {
  const int64_t numel = m * n;
  if (can_vec4(n, {L, Lp, R, Rp, VL, VR}))
    cvx_point_dense_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(L, Lp, R, Rp, numel, beta, t, VL, VR, vr_sumsq);
  else
    cvx_point_dense_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(L, Lp, R, Rp, numel, beta, t, VL, VR, vr_sumsq);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int cvx_finish_dense(const float* W, const float* Lnew, const float* VR, int64_t m, int64_t n, const double* sc,
                     float* Rnew, float* E, cudaStream_t st) {
  const int64_t numel = m * n;
  if (can_vec4(n, {W, Lnew, VR, Rnew, E}))
    cvx_finish_dense_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(W, Lnew, VR, numel, sc, Rnew, E);
  else
    cvx_finish_dense_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(W, Lnew, VR, numel, sc, Rnew, E);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int scale_double(double* x, double f, cudaStream_t st) {
  scale_double_kernel<<<1, 1, 0, st>>>(x, f);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int cvx_shrink(const float* sigma2, int r, float thresh, float tau_star, int constrained, const double* vr_sumsq,
               float t_lambda, float kappa, float q0, float* colw, float* s_out, double* sc, cudaStream_t st) {
  if (r > 512) return CB_ERR_UNSUPPORTED;
  cvx_shrink_kernel<<<1, 512, 0, st>>>(sigma2, r, thresh, tau_star, constrained, vr_sumsq, t_lambda, kappa, q0, colw, s_out, sc);
  CB_CHECK_LAUNCH();
  return CB_OK;
}
int cvx_finish(const float* W, const float* Lnew, const float* VR, const float* h, int64_t m, int64_t n, const double* sc,
               float* Rnew, double* smooth, cudaStream_t st) {
  const int64_t numel = m * n;
  if (can_vec4(n, {W, Lnew, VR, h, Rnew}))
    cvx_finish_kernel<4><<<grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(W, Lnew, VR, h, numel, n, sc, Rnew, smooth);
  else
    cvx_finish_kernel<1><<<grid_for(numel, 256 * 4, 8), 256, 0, st>>>(W, Lnew, VR, h, numel, n, sc, Rnew, smooth);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// requires m % 4 == 0, n % 4 == 0 (the tensor-core path asks for multiples of 8) and 16-byte aligned bases
int form_y_bf16(const float* Ws, const void* codes, int bits, const float* qscale, const float* sqrt_h, int64_t m,
                int64_t n, __nv_bfloat16* Yb, __nv_bfloat16* Ytb, float* RES, cudaStream_t st) {
  if (m % 4 != 0 || n % 4 != 0 || !aligned16(Ws) || !aligned16(Yb) || !aligned16(Ytb)) return CB_ERR_ARG;
  const float lv = (float)((1 << (bits - 1)) - 1);
  dim3 grid((unsigned)((n + 63) / 64), (unsigned)((m + 63) / 64));
  if (bits <= 8)
    form_y_bf16_kernel<int8_t><<<grid, 256, 0, st>>>(Ws, reinterpret_cast<const int8_t*>(codes), qscale, lv, sqrt_h, (int)m, (int)n, Yb, Ytb, RES);
  else
    form_y_bf16_kernel<int16_t><<<grid, 256, 0, st>>>(Ws, reinterpret_cast<const int16_t*>(codes), qscale, lv, sqrt_h, (int)m, (int)n, Yb, Ytb, RES);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

int quant_form_y_bf16(const float* Ws, const float* LR, const float* h_err, const float* sqrt_h, int64_t m, int64_t n,
                      const float* amax, float eps, int bits, void* codes, float* qscale, double* num,
                      __nv_bfloat16* Yb, __nv_bfloat16* Ytb, float* RES, cudaStream_t st) {
  if (m % 4 != 0 || n % 4 != 0 || !aligned16(Ws) || !aligned16(Yb) || !aligned16(Ytb) || (LR != nullptr && !aligned16(LR)))
    return CB_ERR_ARG;
  const float lv = (float)((1 << (bits - 1)) - 1);
  dim3 grid((unsigned)((n + 63) / 64), (unsigned)((m + 63) / 64));
  if (bits <= 8)
    quant_form_y_bf16_kernel<int8_t><<<grid, 256, 0, st>>>(Ws, LR, h_err, sqrt_h, (int)m, (int)n, amax, eps, lv,
                                                           reinterpret_cast<int8_t*>(codes), qscale, num, Yb, Ytb, RES);
  else
    quant_form_y_bf16_kernel<int16_t><<<grid, 256, 0, st>>>(Ws, LR, h_err, sqrt_h, (int)m, (int)n, amax, eps, lv,
                                                            reinterpret_cast<int16_t*>(codes), qscale, num, Yb, Ytb, RES);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// ---------------------------------------------------------------- batched whole-tensor quantiser (LPLR factors)
__global__ void __launch_bounds__(256) absmax_b_kernel(const float* __restrict__ x_, int64_t numel, float* __restrict__ out_, int64_t bstride) {
  __shared__ float red[32];
  const float* __restrict__ x = boff(x_, bstride * blockIdx.y);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float a = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) a = fmaxf(a, fabsf(x[i]));
  a = block_max(a, red);
  if (threadIdx.x == 0) atomic_max_nonneg(boff(out_, bstride * blockIdx.y), a);
}
template <typename code_t>
__global__ void __launch_bounds__(256)
quant_whole_b_kernel(const float* __restrict__ x_, int64_t numel, const float* __restrict__ amax_, float eps, float lv,
                     code_t* __restrict__ codes_, float* __restrict__ scale_out_, float* __restrict__ deq_, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.y;
  const float* __restrict__ x = boff(x_, bo);
  code_t* __restrict__ codes = boff(codes_, bo);
  float* __restrict__ deq = boff(deq_, bo);
  const float s = fmaxf(boff(amax_, bo)[0], eps);                 // quantization.py:262-265
  const ScaleRecip sr = make_scale_recip(s), lvr = make_scale_recip(lv);
  if (blockIdx.x == 0 && threadIdx.x == 0) boff(scale_out_, bo)[0] = s;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const int c = quant_code(x[i], sr, lv);
    codes[i] = (code_t)c;
    if (deq != nullptr) deq[i] = dequant_val(c, s, lvr);
  }
}

int quantize_whole_batched(const float* x, int64_t numel, int bits, void* codes, float* scale_out, float* deq,
                           float* amax_scratch, cudaStream_t st, const Bt& bt) {
  CopySegments z;
  z.add(amax_scratch, nullptr, sizeof(float));
  CB_TRY(copy_if_multi(nullptr, z, st, bt));
  dim3 grid((unsigned)grid_for(numel, 256 * 4, 1), (unsigned)bt.n);
  absmax_b_kernel<<<grid, 256, 0, st>>>(x, numel, amax_scratch, bt.stride);
  CB_CHECK_LAUNCH();
  const float lv = (float)((1 << (bits - 1)) - 1);
  if (bits <= 8)
    quant_whole_b_kernel<int8_t><<<grid, 256, 0, st>>>(x, numel, amax_scratch, 1e-8f, lv, reinterpret_cast<int8_t*>(codes), scale_out, deq, bt.stride);
  else
    quant_whole_b_kernel<int16_t><<<grid, 256, 0, st>>>(x, numel, amax_scratch, 1e-8f, lv, reinterpret_cast<int16_t*>(codes), scale_out, deq, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// ---------------------------------------------------------------- Hadamard pre-rotation
// Normalised Walsh-Hadamard transform (Sylvester ordering, like scipy.linalg.hadamard / sqrt(P)) of every
// row of a matrix: one CTA per row, the row lives in shared memory for the log2(P) butterfly stages.  Rows
// are read from `src` (src_rows x src_cols, zero beyond) so the zero padding of main.py:84-90 is implicit.
__global__ void __launch_bounds__(256)
fwht_rows_kernel(const float* __restrict__ src, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                 float* __restrict__ dst, int64_t P, float norm) {
  extern __shared__ float row[];
  const int64_t r = blockIdx.x;
  if (r >= src_rows) {            // a zero row transforms to a zero row
    for (int64_t j = threadIdx.x; j < P; j += blockDim.x) dst[r * P + j] = 0.f;
    return;
  }
  for (int64_t j = threadIdx.x; j < P; j += blockDim.x) row[j] = j < src_cols ? src[r * src_ld + j] : 0.f;
  __syncthreads();
  for (int64_t len = 1; len < P; len <<= 1) {
    for (int64_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
      const int64_t i = ((t / len) * len << 1) + (t % len), j = i + len;
      const float a = row[i], b = row[j];
      row[i] = a + b;
      row[j] = a - b;
    }
    __syncthreads();
  }
  for (int64_t j = threadIdx.x; j < P; j += blockDim.x) dst[r * P + j] = row[j] * norm;
}

static int fwht_rows(const float* src, int64_t src_rows, int64_t src_cols, int64_t src_ld, float* dst, int64_t nrows,
                     int64_t P, cudaStream_t st) {
  const size_t smem = (size_t)P * sizeof(float);
  // the opt-in is a per-device attribute: always ask for the kernel's maximum (P <= 32768 rows of 4 bytes), once per device
  static PerDeviceOnce once;
  if (smem > 48 * 1024) CB_TRY(opt_in_dynamic_smem(fwht_rows_kernel, 32768 * (int)sizeof(float), once));
  fwht_rows_kernel<<<(unsigned)nrows, 256, smem, st>>>(src, src_rows, src_cols, src_ld, dst, P, 1.f / sqrtf((float)P));
  CB_CHECK_LAUNCH();
  return CB_OK;
}

}  // namespace cb

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" size_t cb_hadamard_workspace_bytes(int64_t prows, int64_t pcols) {
  if (prows <= 0 || pcols <= 0) return 0;
  return (size_t)prows * pcols * sizeof(float) + 256;
}

extern "C" int cb_hadamard_transform_f32(const float* W, int64_t rows, int64_t cols, float* out, int64_t prows,
                                         int64_t pcols, void* ws, size_t ws_bytes, void* stream) {
  using namespace cb;
  if (W == nullptr || out == nullptr || ws == nullptr || rows <= 0 || cols <= 0) return CB_ERR_ARG;
  if (!is_pow2(prows) || !is_pow2(pcols) || prows < rows || pcols < cols) return CB_ERR_ARG;
  if (prows > 32768 || pcols > 32768) return CB_ERR_UNSUPPORTED;      // one row must fit in shared memory
  if (ws_bytes < cb_hadamard_workspace_bytes(prows, pcols)) return CB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* tmp = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  // out = pad(W) H2 (rows of length pcols), then H1 applied to the columns through two transposes
  CB_TRY(fwht_rows(W, rows, cols, cols, out, prows, pcols, st));
  CB_TRY(transpose_codes(out, prows, pcols, 4, tmp, st));                       // tmp: pcols x prows
  CB_TRY(fwht_rows(tmp, pcols, prows, prows, tmp, pcols, prows, st));           // in place, row by row
  CB_TRY(transpose_codes(tmp, pcols, prows, 4, out, st));
  return CB_OK;
}

extern "C" int cb_hessian_probe(const float* H, int64_t n, float* diag, int* is_diag, void* stream) {
  if (H == nullptr || diag == nullptr || is_diag == nullptr || n <= 0) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CB_CUDA(cudaMemsetAsync(is_diag, 0, sizeof(int), st));
  cb::hessian_probe_kernel<<<cb::grid_for(n * n, 256 * 8, 4), 256, 0, st>>>(H, n, diag, is_diag);
  CB_CHECK_LAUNCH();
  cb::probe_finish_kernel<<<1, 1, 0, st>>>(is_diag);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_sum_stats(const float* x, const float* y, int64_t numel, double* out3, void* stream) {
  if (x == nullptr || out3 == nullptr || numel < 0) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cb::stats_kernel<<<cb::grid_for(numel, 256 * 8, 4), 256, 0, st>>>(x, y, numel, out3);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_quantize_residual_f32(const float* R, const float* base, int64_t rows, int64_t cols, int bits,
                                        float* Rq, float* Wc, float* delta_out, float* scratch, void* stream) {
  if (R == nullptr || Rq == nullptr || delta_out == nullptr || scratch == nullptr || rows <= 0 || cols <= 0) return CB_ERR_ARG;
  if (bits < 2 || bits > 16) return CB_ERR_BITS;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t numel = rows * cols;
  CB_TRY(cb::resid_absmax(R, nullptr, numel, scratch, st));
  if (cb::can_vec4(cols, {R, base, Rq, Wc}))
    cb::cvx_quant_residual_kernel<4><<<cb::grid_for(numel / 4, 256 * 2, 8), 256, 0, st>>>(R, base, numel, scratch, bits, Rq, Wc, delta_out);
  else
    cb::cvx_quant_residual_kernel<1><<<cb::grid_for(numel, 256 * 4, 8), 256, 0, st>>>(R, base, numel, scratch, bits, Rq, Wc, delta_out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

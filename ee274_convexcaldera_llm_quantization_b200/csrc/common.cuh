// Shared device/host helpers for libcaldera_b200 (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/caldera_b200.h"

namespace cb {

// Measurement aid, compiled only with -DCB_MEASURE (python -m ...build --measure -> libcaldera_b200_measure.so):
// CB_DEBUG_SKIP is a bit mask of kernel classes whose
// launches are dropped so that their share of a multi-stream run can be read off the change in
// throughput (results are then garbage).  1: Cholesky, 2: Jacobi, 4: tcgen05 contractions.
#ifdef CB_MEASURE
inline int debug_skip() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CB_DEBUG_SKIP"); v = e != nullptr ? atoi(e) : 0; }
  return v;
}
#else
constexpr int debug_skip() { return 0; }     // the release library has no knock-out switch
#endif


// ---------------------------------------------------------------- launch bookkeeping
extern long long g_launch_count;  // defined in api.cu
inline void note_launch(int n = 1) { __atomic_fetch_add(&g_launch_count, (long long)n, __ATOMIC_RELAXED); }

#define CB_CHECK_LAUNCH()                                         \
  do {                                                            \
    cb::note_launch();                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return CB_ERR_CUDA_BASE + (int)e__;   \
  } while (0)

#define CB_CUDA(call)                                             \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return CB_ERR_CUDA_BASE + (int)e__;   \
  } while (0)

#define CB_TRY(call)                   \
  do {                                 \
    int s__ = (call);                  \
    if (s__ != CB_OK) return s__;      \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting and the library is used on several
// devices of one process (caldera(device="cuda:N"), one worker thread per stream): remember, per kernel, which
// devices have been opted in.  Lock-free; a racing second call merely repeats the (idempotent) attribute call.
struct PerDeviceOnce {
  unsigned long long mask[4] = {0ull, 0ull, 0ull, 0ull};     // up to 256 devices
  bool seen(int dev) const {
    return dev >= 0 && dev < 256 && ((__atomic_load_n(&mask[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull) != 0;
  }
  void mark(int dev) {
    if (dev >= 0 && dev < 256) __atomic_fetch_or(&mask[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
  }
};
template <typename Kernel>
inline int opt_in_dynamic_smem(Kernel kernel, int bytes, PerDeviceOnce& once) {
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  if (once.seen(dev)) return CB_OK;
  CB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  once.mark(dev);
  return CB_OK;
}

// Batched launches: layer b of a batch keeps every buffer b * stride bytes after layer 0's (one slab per layer), so
// a batched kernel takes layer-0 pointers plus that stride and picks its layer from a grid dimension.
template <typename T> __host__ __device__ __forceinline__ T* boff(T* p, int64_t bytes) {
  return p == nullptr ? nullptr : reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + (uintptr_t)bytes);
}

constexpr int kNumSMs = 148;  // B200

inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int levels_of(int bits) { return (1 << (bits - 1)) - 1; }

// Bit-exact restatement of the reference arithmetic (quantization.py:266, 95-96):
// IEEE divide, IEEE multiply (no FMA contraction), round-half-even.
__device__ __forceinline__ int quant_code(float x, float s, float lv) {
  return __float2int_rn(__fmul_rn(__fdiv_rn(x, s), lv));
}

// The same correctly rounded quotient x / s with the reciprocal hoisted out of the element
// loop (one scale serves a whole block).  Markstein's theorem: with y = RN(1/s) and q within
// one ulp of x/s, RN(q + (x - s q) y) == RN(x / s) when the remainder is formed exactly
// (FMA), unless the significand of s is all ones.  Two correction steps bring the first
// estimate RN(x y) (< 1.5 ulp) inside that hypothesis.  `ScaleRecip::exact` is false for the
// excluded significand and for scales whose reciprocal or remainders could leave the normal
// range; those (block-uniform, rare) cases take the plain IEEE divide.
struct ScaleRecip {
  float s, y;
  bool exact;
};
__device__ __forceinline__ ScaleRecip make_scale_recip(float s) {
  ScaleRecip r;
  r.s = s;
  r.y = __frcp_rn(s);
  r.exact = (s > 1e-18f) && (s < 1e18f) && ((__float_as_uint(s) & 0x7FFFFFu) != 0x7FFFFFu);
  return r;
}
__device__ __forceinline__ float div_by_scale(float x, const ScaleRecip& r) {
  if (!r.exact) return __fdiv_rn(x, r.s);   // uniform over the block that shares the scale
  // |x| so small that a remainder would be subnormal (inexact) gives |x / s| < 1e-12 here, far
  // below every code boundary, so its last bits cannot change a code
  const float q0 = __fmul_rn(x, r.y);
  const float q1 = fmaf(fmaf(-q0, r.s, x), r.y, q0);
  return fmaf(fmaf(-q1, r.s, x), r.y, q1);
}
__device__ __forceinline__ int quant_code(float x, const ScaleRecip& r, float lv) {
  return __float2int_rn(__fmul_rn(div_by_scale(x, r), lv));
}
// branch-free core of div_by_scale for callers that tested r.exact once per block
__device__ __forceinline__ float div_by_scale_exact(float x, const ScaleRecip& r) {
  const float q0 = __fmul_rn(x, r.y);
  const float q1 = fmaf(fmaf(-q0, r.s, x), r.y, q0);
  return fmaf(fmaf(-q1, r.s, x), r.y, q1);
}
// four codes sharing one scale: the (rare) IEEE-divide fallback is taken once for all four
__device__ __forceinline__ void quant_code4(const float4& v, const ScaleRecip& r, float lv, int& c0, int& c1, int& c2, int& c3) {
  if (r.exact) {
    c0 = __float2int_rn(__fmul_rn(div_by_scale_exact(v.x, r), lv));
    c1 = __float2int_rn(__fmul_rn(div_by_scale_exact(v.y, r), lv));
    c2 = __float2int_rn(__fmul_rn(div_by_scale_exact(v.z, r), lv));
    c3 = __float2int_rn(__fmul_rn(div_by_scale_exact(v.w, r), lv));
  } else {
    c0 = quant_code(v.x, r.s, lv); c1 = quant_code(v.y, r.s, lv);
    c2 = quant_code(v.z, r.s, lv); c3 = quant_code(v.w, r.s, lv);
  }
}
// The ternary ("2-bit", levels = 1) code without any division: rint(fmul(fdiv(x, s), 1)) is +1 exactly when the correctly
// rounded quotient exceeds 0.5 (a tie at 0.5 rounds to the even 0), i.e. when x / s > 0.5 + 2^-25 in real arithmetic;
// s / 2 is exact in fp32 and the next fp32 value above it is s / 2 + ulp(s / 2) >= s / 2 (1 + 2^-23) > s (0.5 + 2^-25),
// so for an fp32 x this is the same as x > s / 2.  Symmetric for -1.  Valid for |x| <= s and a normal s / 2.
__device__ __forceinline__ bool ternary_ok(float s) { return s > 1e-30f && s < 1e30f; }
__device__ __forceinline__ int ternary_code(float x, float half_s) { return (int)(x > half_s) - (int)(x < -half_s); }

// quantization.py:105, 295
__device__ __forceinline__ float dequant_val(int code, float s, float lv) {
  return __fmul_rn(__fdiv_rn((float)code, lv), s);
}
// same value with the reciprocal of the level count hoisted (lv = 2^(b-1) - 1 never has an
// all-ones significand, so the exact path always applies)
__device__ __forceinline__ float dequant_val(int code, float s, const ScaleRecip& lvr) {
  return __fmul_rn(div_by_scale_exact((float)code, lvr), s);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reductions; result valid in thread 0.  `red` is a 32-entry shared array.
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_max(v);
  __syncthreads();
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// max over non-negative floats via integer atomics (bit patterns of x >= 0 are ordered)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// streaming 128-bit loads/stores (data touched once: do not pollute L1)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// counter-based RNG (splitmix64 finaliser) -> standard normal by Box-Muller
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float gaussian_from(uint64_t seed, uint64_t idx) {
  uint64_t b = mix64(seed ^ mix64(idx));
  float u1 = ((float)((b >> 40) & 0xFFFFFF) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = ((float)((b >> 8) & 0xFFFFFF) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace cb

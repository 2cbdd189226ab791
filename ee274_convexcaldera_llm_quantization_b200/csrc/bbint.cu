// bitsandbytes-style asymmetric block quantisers of the reference's QuantizerFactory surface:
// `bbint4` / `bbint2`, RCR/caldera/utils/quantization.py:107-243.  Per block of `block` consecutive elements:
//   mean, unbiased std (clamped at eps); outliers = |x - mean| > 6 std, kept in a side table (values + (row, col) of the
//   blocked view, row-major order) and replaced by the mean; min / max of what is left; scale = max((max - min) / levels,
//   eps) with levels = 15 (4-bit) or 3 (2-bit); code = clamp(rint((x - min) / scale), 0, levels), packed MSB-first.
// One group of threads per block: a warp for blocks up to 2048 elements, a 512-thread CTA above (the whole-tensor
// blocks the reference forces inside caldera(), alg.py:247).  The outlier table needs an ordered compaction: the
// quantising kernel counts outliers per block, a single-CTA scan turns counts into offsets, and a second kernel, which
// only does work in blocks that have outliers, writes them in order.
#include <cstdint>

#include "common.cuh"
#include "internal.h"

namespace cb {

namespace {

struct BlockStats {
  float mean, std6, vmin, scale;
  int outliers;
};

template <bool CTA> __device__ __forceinline__ double group_sum(double v, double* red) {
  if (CTA) { v = block_sum(v, red); __syncthreads(); if (threadIdx.x == 0) red[0] = v; __syncthreads(); v = red[0]; __syncthreads(); return v; }
  v = warp_sum(v);
  return __shfl_sync(0xffffffffu, v, 0);
}
template <bool CTA> __device__ __forceinline__ float group_minmax(float v, bool is_max, float* redf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : fminf(v, t);
  }
  if (!CTA) return v;
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redf[w] = v;
  __syncthreads();
  float r = redf[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, redf[i]) : fminf(r, redf[i]);
  __syncthreads();
  return r;
}
template <bool CTA> __device__ __forceinline__ int group_sum_int(int v, int* redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (!CTA) return v;
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redi[w] = v;
  __syncthreads();
  int r = 0;
  for (int i = 0; i < nw; ++i) r += redi[i];
  __syncthreads();
  return r;
}

// Statistics of one block; `t` / `nt`: this thread's index in its group and the group's size.
template <bool CTA>
__device__ __forceinline__ BlockStats block_stats(const float* __restrict__ x, int64_t block, float eps, float levels,
                                                  int t, int nt, double* red, float* redf, int* redi) {
  BlockStats s;
  double acc = 0.0;
  for (int64_t i = t; i < block; i += nt) acc += (double)x[i];
  s.mean = (float)(group_sum<CTA>(acc, red) / (double)block);              // weight_blocks.mean(dim=1)
  acc = 0.0;
  for (int64_t i = t; i < block; i += nt) { const double d = (double)x[i] - (double)s.mean; acc += d * d; }
  const double var = block > 1 ? group_sum<CTA>(acc, red) / (double)(block - 1) : nan("");   // torch.std: unbiased
  const float sd = fmaxf((float)sqrt(var), eps);                            // (one-element blocks are rejected by the host)
  s.std6 = __fmul_rn(6.0f, sd);
  float vmin = INFINITY, vmax = -INFINITY;
  int cnt = 0;
  for (int64_t i = t; i < block; i += nt) {
    float v = x[i];
    if (fabsf(__fsub_rn(v, s.mean)) > s.std6) { v = s.mean; ++cnt; }       // outliers are replaced by the mean
    vmin = fminf(vmin, v);
    vmax = fmaxf(vmax, v);
  }
  s.vmin = group_minmax<CTA>(vmin, false, redf);
  vmax = group_minmax<CTA>(vmax, true, redf);
  s.outliers = group_sum_int<CTA>(cnt, redi);
  s.scale = fmaxf(__fdiv_rn(__fsub_rn(vmax, s.vmin), levels), eps);
  return s;
}

template <int BITS, bool CTA>
__global__ void __launch_bounds__(CTA ? 512 : 256)
bbint_quant_kernel(const float* __restrict__ x, int64_t nblocks, int64_t block, float eps, uint8_t* __restrict__ packed,
                   float* __restrict__ block_min, float* __restrict__ scales, int* __restrict__ counts) {
  __shared__ double red[32];
  __shared__ float redf[32];
  __shared__ int redi[32];
  constexpr int E = 8 / BITS;                  // elements per packed byte
  constexpr float LEVELS = (float)((1 << BITS) - 1);
  const int nt = CTA ? blockDim.x : 32;
  const int t = CTA ? threadIdx.x : (threadIdx.x & 31);
  const int64_t groups = CTA ? gridDim.x : ((int64_t)gridDim.x * (blockDim.x >> 5));
  const int64_t g0 = CTA ? blockIdx.x : ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  for (int64_t b = g0; b < nblocks; b += groups) {
    const float* xb = x + b * block;
    const BlockStats s = block_stats<CTA>(xb, block, eps, LEVELS, t, nt, red, redf, redi);
    if (t == 0) { block_min[b] = s.vmin; scales[b] = s.scale; counts[b] = s.outliers; }
    uint8_t* pb = packed + b * (block / E);
    for (int64_t k = t; k < block / E; k += nt) {
      uint32_t byte = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float v = xb[k * E + e];
        if (fabsf(__fsub_rn(v, s.mean)) > s.std6) v = s.mean;
        float q = rintf(__fdiv_rn(__fsub_rn(v, s.vmin), s.scale));          // torch.round: half to even
        q = fminf(fmaxf(q, 0.f), LEVELS);
        byte = (byte << BITS) | (uint32_t)q;                                // element 0 in the most significant bits
      }
      pb[k] = (uint8_t)byte;
    }
  }
}

// exclusive scan of counts[0..n) into offsets[0..n), total into offsets[n]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) bbint_scan_kernel(const int* __restrict__ counts, int64_t n, int64_t* __restrict__ offsets) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n ? (int64_t)counts[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t tt = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += tt;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      int64_t ws = wsum[lane], winc = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t tt = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += tt;
      }
      wsum[lane] = winc - ws;                   // exclusive prefix of the warp sums
    }
    __syncthreads();
    const int64_t excl = carry + wsum[w] + inc - v;
    if (i < n) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// The side table: for every block with outliers (counts[b] > 0), its outliers in element order.
template <bool CTA>
__global__ void __launch_bounds__(CTA ? 512 : 256)
bbint_outliers_kernel(const float* __restrict__ x, int64_t nblocks, int64_t block, float eps, const int* __restrict__ counts,
                      const int64_t* __restrict__ offsets, float* __restrict__ values, int64_t* __restrict__ indices) {
  __shared__ double red[32];
  __shared__ float redf[32];
  __shared__ int redi[32];
  __shared__ int wcount[32];
  const int nt = CTA ? blockDim.x : 32;
  const int t = CTA ? threadIdx.x : (threadIdx.x & 31);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t groups = CTA ? gridDim.x : ((int64_t)gridDim.x * (blockDim.x >> 5));
  const int64_t g0 = CTA ? blockIdx.x : ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  for (int64_t b = g0; b < nblocks; b += groups) {
    if (counts[b] == 0) continue;               // uniform over the group
    const float* xb = x + b * block;
    const BlockStats s = block_stats<CTA>(xb, block, eps, 1.f, t, nt, red, redf, redi);
    int64_t pos = offsets[b];
    for (int64_t base = 0; base < block; base += nt) {
      const int64_t i = base + t;
      const float v = i < block ? xb[i] : 0.f;
      const bool out = i < block && fabsf(__fsub_rn(v, s.mean)) > s.std6;
      const unsigned bal = __ballot_sync(0xffffffffu, out);
      int before = __popc(bal & ((1u << lane) - 1u));
      int chunk_total = __popc(bal);
      if (CTA) {
        __syncthreads();
        if (lane == 0) wcount[w] = chunk_total;
        __syncthreads();
        int pre = 0, tot = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { if (k < w) pre += wcount[k]; tot += wcount[k]; }
        before += pre;
        chunk_total = tot;
      }
      if (out) {
        values[pos + before] = v;
        indices[2 * (pos + before)] = b;
        indices[2 * (pos + before) + 1] = i;
      }
      pos += chunk_total;
    }
  }
}

template <int BITS>
__global__ void __launch_bounds__(256)
bbint_dequant_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ block_min, const float* __restrict__ scales,
                     int64_t numel, int64_t block, float* __restrict__ out) {
  constexpr int E = 8 / BITS;
  const int64_t nbytes = numel / E, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nbytes; k += stride) {
    const uint32_t byte = packed[k];
    const int64_t i0 = k * E, b = i0 / block;
    const float sc = scales[b], mn = block_min[b];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float q = (float)((byte >> (BITS * (E - 1 - e))) & ((1u << BITS) - 1u));
      out[i0 + e] = __fadd_rn(__fmul_rn(q, sc), mn);                        // weight_unpacked * scales + block_min
    }
  }
}

__global__ void __launch_bounds__(256)
bbint_restore_kernel(const float* __restrict__ values, const int64_t* __restrict__ indices, int64_t count, int64_t block,
                     int64_t numel, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const int64_t pos = indices[2 * k] * block + indices[2 * k + 1];
    if (pos >= 0 && pos < numel) out[pos] = values[k];
  }
}

}  // namespace

}  // namespace cb

using namespace cb;

static bool bbint_args_ok(int64_t numel, int64_t block, int bits) {
  return (bits == 2 || bits == 4) && numel > 0 && block >= 2 && numel % block == 0 && block % (8 / bits) == 0;
}

extern "C" int cb_quantize_bbint_f32(const float* x, int64_t numel, int64_t block, int bits, float eps, uint8_t* packed,
                                     float* block_min, float* scales, int* counts, int64_t* offsets, void* stream) {
  if (x == nullptr || packed == nullptr || block_min == nullptr || scales == nullptr || counts == nullptr || offsets == nullptr)
    return CB_ERR_ARG;
  if (!bbint_args_ok(numel, block, bits)) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = numel / block;
  if (block <= 2048) {
    const int grid = grid_for(nblocks, 8, 16);
    if (bits == 4) bbint_quant_kernel<4, false><<<grid, 256, 0, st>>>(x, nblocks, block, eps, packed, block_min, scales, counts);
    else bbint_quant_kernel<2, false><<<grid, 256, 0, st>>>(x, nblocks, block, eps, packed, block_min, scales, counts);
  } else {
    const int grid = grid_for(nblocks, 1, 4);
    if (bits == 4) bbint_quant_kernel<4, true><<<grid, 512, 0, st>>>(x, nblocks, block, eps, packed, block_min, scales, counts);
    else bbint_quant_kernel<2, true><<<grid, 512, 0, st>>>(x, nblocks, block, eps, packed, block_min, scales, counts);
  }
  CB_CHECK_LAUNCH();
  bbint_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, offsets);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_bbint_outliers_f32(const float* x, int64_t numel, int64_t block, float eps, const int* counts,
                                     const int64_t* offsets, float* values, int64_t* indices, void* stream) {
  if (x == nullptr || counts == nullptr || offsets == nullptr || values == nullptr || indices == nullptr) return CB_ERR_ARG;
  if (numel <= 0 || block < 2 || numel % block != 0) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = numel / block;
  if (block <= 2048)
    bbint_outliers_kernel<false><<<grid_for(nblocks, 8, 16), 256, 0, st>>>(x, nblocks, block, eps, counts, offsets, values, indices);
  else
    bbint_outliers_kernel<true><<<grid_for(nblocks, 1, 4), 512, 0, st>>>(x, nblocks, block, eps, counts, offsets, values, indices);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_dequantize_bbint_f32(const uint8_t* packed, const float* block_min, const float* scales, int64_t numel,
                                       int64_t block, int bits, const float* outlier_values, const int64_t* outlier_indices,
                                       int64_t n_outliers, float* out, void* stream) {
  if (packed == nullptr || block_min == nullptr || scales == nullptr || out == nullptr) return CB_ERR_ARG;
  if (!bbint_args_ok(numel, block, bits) || n_outliers < 0) return CB_ERR_ARG;
  if (n_outliers > 0 && (outlier_values == nullptr || outlier_indices == nullptr)) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nbytes = numel / (8 / bits);
  if (bits == 4) bbint_dequant_kernel<4><<<grid_for(nbytes, 256 * 4, 8), 256, 0, st>>>(packed, block_min, scales, numel, block, out);
  else bbint_dequant_kernel<2><<<grid_for(nbytes, 256 * 4, 8), 256, 0, st>>>(packed, block_min, scales, numel, block, out);
  CB_CHECK_LAUNCH();
  if (n_outliers > 0) {
    bbint_restore_kernel<<<grid_for(n_outliers, 256, 4), 256, 0, st>>>(outlier_values, outlier_indices, n_outliers, block, numel, out);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

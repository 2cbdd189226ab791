// Block-wise abs-max uniform quantiser, dequantiser and the packed code format.
//
// Replaces LowMemoryQuantizer.quantize_block / dequantize_block (uniform branch),
// RCR/caldera/utils/quantization.py:244-268, 290-307.  HBM-bound: every kernel here reads
// its input once with 128-bit streaming loads, keeps the per-block abs-max / scale in
// registers (warp-shuffle reduction over the lanes that share a block) and writes codes,
// packed codes, scales and (optionally) the dequantised values from the same registers.
//
// Algorithmic bytes per element ("quantise/pack"): 4 (read) + bits/8 (packed)
// + 4/block (scales) [+1 if int8 codes are requested, +4 if dequantised output is].
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cb {

// ---------------------------------------------------------------- packing helpers
template <int BITS> struct CodeT { using type = int8_t; };
template <> struct CodeT<16> { using type = int16_t; };

// ---------------------------------------------------------------- fast path
// One warp handles tiles of 1024 consecutive elements as 4 chunks of 256: every lane issues
// one 256-bit streaming load per chunk (LDG.E.256, 1 KiB per warp instruction, all four in
// flight before the first use) and owns 8 consecutive elements of it.  LPB = lanes that share
// a quantisation block (block / 8) for block in {32, 64, 128, 256}; the block abs-max is a
// log2(LPB)-step shuffle reduction and never leaves registers.  LPB == 0 means one scale for
// the whole tensor, already reduced into scales[0] by absmax_fast_kernel.  FULL: numel is a
// multiple of 1024, so no lane is ever out of range.
struct F8 { float v[8]; };

__device__ __forceinline__ F8 ld_stream8(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

// 1.0f / 0.0f comparison results (FSET.BF): a compare and a multiply-add put a symbol bit into place
__device__ __forceinline__ float fset_gt(float a, float b) {
  float d;
  asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fset_geu(float a, float b) {
  float d;
  asm("set.geu.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// One tile (4 chunks of 256 elements, lane owns 8 consecutive elements of each) from registers to codes / packed codes /
// scales / dequantised values.  Shared by the register-staged and the shared-memory-staged kernel below.
//
// The four chunks advance through the phases together (abs-max, shuffle reduction, scale, codes, stores), so that a warp
// has four independent dependency chains in flight instead of one.  PACK_ONLY (only bit-packed codes and scales are
// wanted: the "quantise + pack" operation of the wire format) with 2- or 4-bit codes takes a branch-free path whenever
// every scale of the warp's tile admits it (a warp vote; anything else falls back to the general per-chunk code):
//   2-bit: the packed symbol code + 1 = (x > s/2) + !(x < -s/2) is accumulated straight into its bit position, two
//          compare-and-add pairs per element and no separate packing step;
//   4-bit: correctly rounded reciprocal-multiply codes, shifted into place by one multiply-add each.
template <int BITS, int LPB, bool PACK_ONLY>
__device__ __forceinline__ void quant_tile(const F8 (&v)[4], const bool (&ok)[4], const int64_t base, const int lane,
                                           const float lv, const ScaleRecip& lvr, const ScaleRecip& whole, const float eps,
                                           void* __restrict__ codes, uint8_t* __restrict__ packed,
                                           float* __restrict__ scales, float* __restrict__ dequant) {
  using code_t = typename CodeT<BITS>::type;
  constexpr int LV = (1 << (BITS - 1)) - 1;
  constexpr int BLOCK = LPB * 8;            // elements per quantisation block (compile time)
  float sc[4];
  if (LPB > 0) {
    float a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j] = fabsf(v[j].v[0]);
#pragma unroll
      for (int e = 1; e < 8; ++e) a[j] = fmaxf(a[j], fabsf(v[j].v[e]));
    }
#pragma unroll
    for (int o = LPB / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = fmaxf(a[j], __shfl_xor_sync(0xffffffffu, a[j], o));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc[j] = fmaxf(a[j], eps);
      if (ok[j] && (lane % (LPB > 0 ? LPB : 1)) == 0) scales[(base + j * 256) / (BLOCK > 0 ? BLOCK : 1)] = sc[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[j] = whole.s;
  }
  if (PACK_ONLY && BITS == 2) {
    const bool tern = ternary_ok(sc[0]) && ternary_ok(sc[1]) && ternary_ok(sc[2]) && ternary_ok(sc[3]);
    if (__all_sync(0xffffffffu, tern)) {
      uint32_t h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hs = 0.5f * sc[j], nhs = -hs;
        float acc = 0.f;                   // exact: the sum stays below 2^16
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // byte = q0*64 + q1*16 + q2*4 + q3 (quantization.py:217-220); elements 4..7 form the lane's second byte
          const float unit = (float)(1u << (e < 4 ? 6 - 2 * e : 14 - 2 * (e - 4)));
          acc = fmaf(fset_gt(v[j].v[e], hs), unit, acc);        // x > s/2
          acc = fmaf(fset_geu(v[j].v[e], nhs), unit, acc);      // !(x < -s/2): a NaN keeps the reference's code 0
        }
        h[j] = __float2uint_rz(acc);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = h[j] | (__shfl_down_sync(0xffffffffu, h[j], 1) << 16);
        if (ok[j] && (lane & 1) == 0) *reinterpret_cast<uint32_t*>(packed + ((base + j * 256) >> 2)) = w;
      }
      return;
    }
  }
  if (PACK_ONLY && BITS == 4) {
    ScaleRecip r4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r4[j] = LPB > 0 ? make_scale_recip(sc[j]) : whole;
    if (__all_sync(0xffffffffu, r4[0].exact && r4[1].exact && r4[2].exact && r4[3].exact)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // byte = q0*16 + q1 (quantization.py:152): four bytes per lane, little-endian in the 32-bit word
        int acc = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int sh = 8 * (e >> 1) + ((e & 1) ? 0 : 4);
          acc += (__float2int_rn(__fmul_rn(div_by_scale_exact(v[j].v[e], r4[j]), lv)) + LV) << sh;
        }
        if (ok[j]) *reinterpret_cast<uint32_t*>(packed + ((base + j * 256) >> 1)) = (uint32_t)acc;
      }
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t idx = base + j * 256;
    const ScaleRecip sr = LPB > 0 ? make_scale_recip(sc[j]) : whole;
    const float s = sr.s;
    int c[8];
    if (BITS == 2 && ternary_ok(s)) {     // two compares per element instead of an exact division (common.cuh)
      const float hs = 0.5f * s;
#pragma unroll
      for (int e = 0; e < 8; ++e) c[e] = ternary_code(v[j].v[e], hs);
    } else if (sr.exact) {
#pragma unroll
      for (int e = 0; e < 8; ++e) c[e] = __float2int_rn(__fmul_rn(div_by_scale_exact(v[j].v[e], sr), lv));
    } else {   // uniform over the lanes that share the scale, rare
#pragma unroll
      for (int e = 0; e < 8; ++e) c[e] = quant_code(v[j].v[e], s, lv);
    }
    if (!PACK_ONLY && codes != nullptr && ok[j]) {
      if (BITS <= 8) {
        uint2 w;
        w.x = (uint32_t)(uint8_t)(int8_t)c[0] | ((uint32_t)(uint8_t)(int8_t)c[1] << 8) |
              ((uint32_t)(uint8_t)(int8_t)c[2] << 16) | ((uint32_t)(uint8_t)(int8_t)c[3] << 24);
        w.y = (uint32_t)(uint8_t)(int8_t)c[4] | ((uint32_t)(uint8_t)(int8_t)c[5] << 8) |
              ((uint32_t)(uint8_t)(int8_t)c[6] << 16) | ((uint32_t)(uint8_t)(int8_t)c[7] << 24);
        *reinterpret_cast<uint2*>(reinterpret_cast<int8_t*>(codes) + idx) = w;
      } else {
        uint4 w;
        w.x = (uint32_t)(uint16_t)(int16_t)c[0] | ((uint32_t)(uint16_t)(int16_t)c[1] << 16);
        w.y = (uint32_t)(uint16_t)(int16_t)c[2] | ((uint32_t)(uint16_t)(int16_t)c[3] << 16);
        w.z = (uint32_t)(uint16_t)(int16_t)c[4] | ((uint32_t)(uint16_t)(int16_t)c[5] << 16);
        w.w = (uint32_t)(uint16_t)(int16_t)c[6] | ((uint32_t)(uint16_t)(int16_t)c[7] << 16);
        *reinterpret_cast<uint4*>(reinterpret_cast<code_t*>(codes) + idx) = w;
      }
    }
    if (!PACK_ONLY && dequant != nullptr && ok[j]) {
      st_stream4(dequant + idx, make_float4(dequant_val(c[0], s, lvr), dequant_val(c[1], s, lvr),
                                            dequant_val(c[2], s, lvr), dequant_val(c[3], s, lvr)));
      st_stream4(dequant + idx + 4, make_float4(dequant_val(c[4], s, lvr), dequant_val(c[5], s, lvr),
                                                dequant_val(c[6], s, lvr), dequant_val(c[7], s, lvr)));
    }
    if (packed != nullptr) {
      if (BITS == 2) {
        // byte = q0*64 + q1*16 + q2*4 + q3 (quantization.py:217-220): two bytes per lane, two
        // neighbouring lanes are gathered into one 32-bit store
        const uint32_t b0 = (uint32_t)(((c[0] + LV) << 6) | ((c[1] + LV) << 4) | ((c[2] + LV) << 2) | (c[3] + LV));
        const uint32_t b1 = (uint32_t)(((c[4] + LV) << 6) | ((c[5] + LV) << 4) | ((c[6] + LV) << 2) | (c[7] + LV));
        const uint32_t h = b0 | (b1 << 8);
        const uint32_t w = h | (__shfl_down_sync(0xffffffffu, h, 1) << 16);
        if (ok[j] && (lane & 1) == 0) *reinterpret_cast<uint32_t*>(packed + (idx >> 2)) = w;
      } else if (BITS == 4) {
        // byte = q0*16 + q1 (quantization.py:152): four bytes per lane
        const uint32_t w = (uint32_t)(((c[0] + LV) << 4) | (c[1] + LV)) | ((uint32_t)(((c[2] + LV) << 4) | (c[3] + LV)) << 8) |
                           ((uint32_t)(((c[4] + LV) << 4) | (c[5] + LV)) << 16) | ((uint32_t)(((c[6] + LV) << 4) | (c[7] + LV)) << 24);
        if (ok[j]) *reinterpret_cast<uint32_t*>(packed + (idx >> 1)) = w;
      } else if (BITS == 8) {
        uint2 w;
        w.x = (uint32_t)(c[0] + LV) | ((uint32_t)(c[1] + LV) << 8) | ((uint32_t)(c[2] + LV) << 16) | ((uint32_t)(c[3] + LV) << 24);
        w.y = (uint32_t)(c[4] + LV) | ((uint32_t)(c[5] + LV) << 8) | ((uint32_t)(c[6] + LV) << 16) | ((uint32_t)(c[7] + LV) << 24);
        if (ok[j]) *reinterpret_cast<uint2*>(packed + idx) = w;
      } else {
        // big-endian offset uint16
        uint32_t sw[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { const uint32_t sy = (uint32_t)(c[e] + LV); sw[e] = (sy >> 8) | ((sy & 255u) << 8); }
        uint4 w;
        w.x = sw[0] | (sw[1] << 16); w.y = sw[2] | (sw[3] << 16); w.z = sw[4] | (sw[5] << 16); w.w = sw[6] | (sw[7] << 16);
        if (ok[j]) *reinterpret_cast<uint4*>(packed + idx * 2) = w;
      }
    }
  }
}

template <int BITS, int LPB, bool FULL, bool PACK_ONLY>
__global__ void __launch_bounds__(256)
quant_fast_kernel(const float* __restrict__ x, int64_t numel, float eps,
                  void* __restrict__ codes, uint8_t* __restrict__ packed,
                  float* __restrict__ scales, float* __restrict__ dequant) {
  constexpr int LV = (1 << (BITS - 1)) - 1;
  const float lv = (float)LV;
  const ScaleRecip lvr = make_scale_recip(lv);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ntiles = (numel + 1023) >> 10;
  ScaleRecip whole = make_scale_recip(1.f);
  if (LPB == 0) whole = make_scale_recip(fmaxf(scales[0], eps));

  for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
    const int64_t base = (tile << 10) + lane * 8;
    F8 v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t idx = base + j * 256;
      ok[j] = FULL || idx < numel;
      if (ok[j]) v[j] = ld_stream8(x + idx);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j].v[e] = 0.f;
      }
    }
    quant_tile<BITS, LPB, PACK_ONLY>(v, ok, base, lane, lv, lvr, whole, eps, codes, packed, scales, dequant);
  }
}

// ---------------------------------------------------------------- shared-memory-staged path (large tensors)
// Same arithmetic, but the input reaches the SM through bulk asynchronous copies (cp.async.bulk, the TMA engine's 1-D
// form) into per-warp rings of 4 KiB tiles in shared memory: the bytes in flight (warps x stages x 4 KiB per SM, ~190 KiB)
// no longer depend on registers or on how many warps are between a load and its use, one CTA per SM stays resident for
// the whole tensor (no second wave, no per-CTA ramp) and every warp computes on tile t while tiles t+1 .. t+stages-1
// are on their way.  A warp owns its ring: lane 0 re-arms a stage's mbarrier and issues the next copy into it after the
// whole warp has consumed the stage, so there is no producer warp and no "empty" barrier.  numel % 1024 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int QS_MAX_THREADS = 768;
constexpr int QS_TILE_BYTES = 4096;

template <int BITS, int LPB, bool PACK_ONLY>
__global__ void __launch_bounds__(QS_MAX_THREADS, 1)
quant_stream_kernel(const float* __restrict__ x, int64_t ntiles, int stages, float eps,
                    void* __restrict__ codes, uint8_t* __restrict__ packed,
                    float* __restrict__ scales, float* __restrict__ dequant) {
  extern __shared__ __align__(128) uint8_t qs_smem[];
  constexpr int LV = (1 << (BITS - 1)) - 1;
  const float lv = (float)LV;
  const ScaleRecip lvr = make_scale_recip(lv);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t ring = smem_u32(qs_smem) + (uint32_t)(warp * stages) * QS_TILE_BYTES;
  const uint32_t bars = smem_u32(qs_smem) + (uint32_t)(nwarp * stages) * QS_TILE_BYTES + (uint32_t)(warp * stages) * 8u;
  const float* ring_g = reinterpret_cast<const float*>(qs_smem + (size_t)(warp * stages) * QS_TILE_BYTES);
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();
  // a launch with programmatic stream serialisation may start while its predecessor drains: let the successor do the
  // same, and touch no global memory before the predecessor has completed
  grid_dependency_launch();
  grid_dependency_wait();

  const int64_t gw = (int64_t)blockIdx.x * nwarp + warp, nW = (int64_t)gridDim.x * nwarp;
  ScaleRecip whole = make_scale_recip(1.f);
  if (LPB == 0) whole = make_scale_recip(fmaxf(scales[0], eps));
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) {
      const int64_t t = gw + (int64_t)s * nW;
      if (t < ntiles) {
        mbar_expect_tx(bars + 8u * s, QS_TILE_BYTES);
        bulk_g2s(ring + (uint32_t)s * QS_TILE_BYTES, x + (t << 10), QS_TILE_BYTES, bars + 8u * s);
      }
    }
  }
  int s = 0;
  uint32_t phase = 0;
  const bool ok[4] = {true, true, true, true};
  for (int64_t t = gw; t < ntiles; t += nW) {
    if (!mbar_wait(bars + 8u * s, phase)) __trap();   // bounded: a protocol bug fails the launch instead of hanging
    F8 v[4];
    const float* src = ring_g + s * (QS_TILE_BYTES / 4) + lane * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(src + j * 256);
      const float4 b = *reinterpret_cast<const float4*>(src + j * 256 + 4);
      v[j].v[0] = a.x; v[j].v[1] = a.y; v[j].v[2] = a.z; v[j].v[3] = a.w;
      v[j].v[4] = b.x; v[j].v[5] = b.y; v[j].v[6] = b.z; v[j].v[7] = b.w;
    }
    quant_tile<BITS, LPB, PACK_ONLY>(v, ok, (t << 10) + lane * 8, lane, lv, lvr, whole, eps, codes, packed, scales, dequant);
    // every lane has used what it read from the stage (its stores depend on it): refill the stage
    __syncwarp();
    if (lane == 0) {
      const int64_t tn = t + (int64_t)stages * nW;
      if (tn < ntiles) {
        mbar_expect_tx(bars + 8u * s, QS_TILE_BYTES);
        bulk_g2s(ring + (uint32_t)s * QS_TILE_BYTES, x + (tn << 10), QS_TILE_BYTES, bars + 8u * s);
      }
    }
    if (++s == stages) { s = 0; phase ^= 1u; }
  }
}

// whole-tensor abs-max, contiguous input (32-byte aligned, numel % 8 == 0 for the vector part); result atomically
// max-ed into *out.  Every thread keeps up to eight independent 256-bit loads in flight (predicated, so that the last
// partial round is issued together as well instead of one load at a time).
__global__ void __launch_bounds__(256)
absmax_fast_kernel(const float* __restrict__ x, int64_t numel, float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f;
  const int64_t n8 = numel >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += 8 * stride) {
    F8 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t idx = i + k * stride;
      if (idx < n8) v[k] = ld_stream8(x + 8 * idx);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k].v[e] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaxf(a, fabsf(v[k].v[e]));
    }
  }
  for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) a = fmaxf(a, fabsf(x[i]));
  a = block_max(a, red);
  if (threadIdx.x == 0) atomic_max_nonneg(out, a);
}

// ---------------------------------------------------------------- generic path
// Any block size, any strides, any numel.  Scales must be zeroed first.
__global__ void __launch_bounds__(256)
absmax_generic_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t sr, int64_t sc,
                      int64_t block, float* __restrict__ scales) {
  const int64_t numel = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < numel; i0 += stride) {
    const int64_t i = i0 + lane;
    float a = 0.f;
    int64_t b = -1;
    if (i < numel) {
      int64_t r = i / cols, c = i - r * cols;
      a = fabsf(x[r * sr + c * sc]);
      b = i / block;
    }
    // warp-aggregate when the whole warp falls into one block
    const int64_t b_first = __shfl_sync(0xffffffffu, b, 0);
    const bool same = __all_sync(0xffffffffu, b == b_first || b < 0);
    if (same) {
      a = warp_max(a);
      if (lane == 0 && b_first >= 0) atomic_max_nonneg(scales + b_first, a);
    } else if (b >= 0) {
      atomic_max_nonneg(scales + b, a);
    }
  }
}

template <int BITS>
__global__ void __launch_bounds__(256)
quant_generic_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t sr, int64_t sc,
                     int64_t block, float eps, void* __restrict__ codes, uint8_t* __restrict__ packed,
                     float* __restrict__ scales, float* __restrict__ dequant) {
  using code_t = typename CodeT<BITS>::type;
  constexpr int LV = (1 << (BITS - 1)) - 1;
  constexpr int G = BITS == 2 ? 4 : (BITS == 4 ? 2 : 1);  // elements per packed byte (or per thread)
  const float lv = (float)LV;
  const int64_t numel = rows * cols;
  const int64_t ngroups = (numel + G - 1) / G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    uint32_t byte = 0;
#pragma unroll
    for (int e = 0; e < G; ++e) {
      const int64_t i = g * G + e;
      int sym = 0;
      if (i < numel) {
        const int64_t r = i / cols, c = i - r * cols;
        const int64_t b = i / block;
        const float s = fmaxf(scales[b], eps);
        if (i == b * block) scales[b] = s;  // idempotent clamp (max(max(a,eps),eps) == max(a,eps))
        const int code = quant_code(x[r * sr + c * sc], s, lv);
        if (codes != nullptr) reinterpret_cast<code_t*>(codes)[i] = (code_t)code;
        if (dequant != nullptr) dequant[i] = dequant_val(code, s, lv);
        sym = code + LV;
      }
      if (BITS <= 4) byte = (byte << BITS) | (uint32_t)sym;
      else byte = (uint32_t)sym;
    }
    if (packed != nullptr) {
      if (BITS <= 8) packed[g] = (uint8_t)byte;
      else { packed[2 * g] = (uint8_t)(byte >> 8); packed[2 * g + 1] = (uint8_t)(byte & 255u); }
    }
  }
}

// ---------------------------------------------------------------- dequantise
template <int BITS, bool PACKED>
__device__ __forceinline__ int fetch_code(const void* codes, const uint8_t* packed, int64_t i) {
  constexpr int LV = (1 << (BITS - 1)) - 1;
  if (!PACKED) return (int)reinterpret_cast<const typename CodeT<BITS>::type*>(codes)[i];
  if (BITS == 2) return (int)((packed[i >> 2] >> (2 * (3 - (int)(i & 3)))) & 3u) - LV;
  if (BITS == 4) return (int)((packed[i >> 1] >> (4 * (1 - (int)(i & 1)))) & 15u) - LV;
  if (BITS == 8) return (int)packed[i] - LV;
  return (int)(((uint32_t)packed[2 * i] << 8) | packed[2 * i + 1]) - LV;
}

template <int BITS, bool PACKED>
__global__ void __launch_bounds__(256)
dequant_generic_kernel(const void* __restrict__ codes, const uint8_t* __restrict__ packed,
                       const float* __restrict__ scales, int64_t numel, int64_t block,
                       float* __restrict__ out) {
  const float lv = (float)((1 << (BITS - 1)) - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    out[i] = dequant_val(fetch_code<BITS, PACKED>(codes, packed, i), scales[i / block], lv);
}

// 16 elements per thread: one 32-bit word of 2-bit codes / two words of 4-bit codes /
// one 128-bit word of int8 codes -> four 128-bit stores.  numel % 16 == 0, block % 16 == 0.
template <int BITS, bool PACKED>
__global__ void __launch_bounds__(256)
dequant_fast_kernel(const void* __restrict__ codes, const uint8_t* __restrict__ packed,
                    const float* __restrict__ scales, int64_t numel, int64_t block,
                    float* __restrict__ out) {
  constexpr int LV = (1 << (BITS - 1)) - 1;
  const ScaleRecip lv = make_scale_recip((float)LV);
  const int64_t n16 = numel >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += stride) {
    const int64_t i = t << 4;
    const float s = scales[i / block];
    int c[16];
    if (PACKED && BITS == 2) {
      uint32_t w = *reinterpret_cast<const uint32_t*>(packed + (i >> 2));
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint32_t byte = (w >> (8 * b)) & 255u;
#pragma unroll
        for (int e = 0; e < 4; ++e) c[4 * b + e] = (int)((byte >> (2 * (3 - e))) & 3u) - LV;
      }
    } else if (PACKED && BITS == 4) {
      uint2 w = *reinterpret_cast<const uint2*>(packed + (i >> 1));
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        uint32_t byte = ((b < 4 ? w.x : w.y) >> (8 * (b & 3))) & 255u;
        c[2 * b] = (int)(byte >> 4) - LV;
        c[2 * b + 1] = (int)(byte & 15u) - LV;
      }
    } else {
      // int8 codes (or offset bytes when PACKED && BITS == 8)
      uint4 w = *reinterpret_cast<const uint4*>(PACKED ? (const void*)(packed + i)
                                                       : (const void*)(reinterpret_cast<const int8_t*>(codes) + i));
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        uint32_t byte = (ww[b >> 2] >> (8 * (b & 3))) & 255u;
        c[b] = PACKED ? (int)byte - LV : (int)(int8_t)byte;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      st_stream4(out + i + 4 * k, make_float4(dequant_val(c[4 * k], s, lv), dequant_val(c[4 * k + 1], s, lv),
                                              dequant_val(c[4 * k + 2], s, lv), dequant_val(c[4 * k + 3], s, lv)));
  }
}

template <int BITS>
__global__ void __launch_bounds__(256)
pack_kernel(const void* __restrict__ codes, int64_t numel, uint8_t* __restrict__ packed) {
  using code_t = typename CodeT<BITS>::type;
  constexpr int LV = (1 << (BITS - 1)) - 1;
  constexpr int G = BITS == 2 ? 4 : (BITS == 4 ? 2 : 1);
  const int64_t ngroups = (numel + G - 1) / G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const code_t* c = reinterpret_cast<const code_t*>(codes);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    uint32_t byte = 0;
#pragma unroll
    for (int e = 0; e < G; ++e) {
      const int64_t i = g * G + e;
      const int sym = i < numel ? (int)c[i] + LV : 0;
      if (BITS <= 4) byte = (byte << BITS) | (uint32_t)sym;
      else byte = (uint32_t)sym;
    }
    if (BITS <= 8) packed[g] = (uint8_t)byte;
    else { packed[2 * g] = (uint8_t)(byte >> 8); packed[2 * g + 1] = (uint8_t)(byte & 255u); }
  }
}

template <int BITS>
__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t* __restrict__ packed, int64_t numel, void* __restrict__ codes) {
  using code_t = typename CodeT<BITS>::type;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    reinterpret_cast<code_t*>(codes)[i] = (code_t)fetch_code<BITS, true>(nullptr, packed, i);
}

// ---------------------------------------------------------------- host dispatch
// Geometry of quant_stream_kernel: warps per CTA x ring stages per warp (4 KiB each) and whether the launch carries the
// programmatic-stream-serialisation attribute (its CTAs are scheduled while the previous kernel of the stream drains
// and wait in griddepcontrol.wait until that kernel has completed).  Measured at 4096 x 4096 on B200
// (scripts/probe_quant_stream.py, profiles/logs/u3_quant_stream.log): 2-bit pack-only 13.5 us with 16 warps x 2 stages and
// the attribute (14.9 without it, 14.4 for the register-staged kernel); the 4- and 8-bit kernels, whose arithmetic
// (correctly rounded reciprocal-multiply) needs the 32 resident warps of the register-staged kernel to hide its
// latency, are faster there (15.1 against 16.2 us), so only the 2-bit packer takes the shared-memory-staged kernel.
// The -DCB_MEASURE build reads CB_QS_WARPS / CB_QS_STAGES / CB_QS_PDL so that the probe can sweep them (warps = 0: never).
struct QuantStreamPolicy { int warps, stages, pdl; };
#ifdef CB_MEASURE
constexpr bool kStreamEveryWidth = true;     // the probe may send any width through quant_stream_kernel
#else
constexpr bool kStreamEveryWidth = false;    // release: only the instantiations the policy below can select are compiled
#endif
constexpr int QS_MAX_SMEM = 216 * 1024;
static QuantStreamPolicy quant_stream_policy(int bits, bool pack_only) {
  QuantStreamPolicy p = {(bits == 2 && pack_only) ? 16 : 0, 2, 1};
#ifdef CB_MEASURE
  if (const char* e = getenv("CB_QS_WARPS")) p.warps = atoi(e);
  if (const char* e = getenv("CB_QS_STAGES")) p.stages = atoi(e);
  if (const char* e = getenv("CB_QS_PDL")) p.pdl = atoi(e);
  if (p.warps < 0 || p.warps > QS_MAX_THREADS / 32 || p.stages < 1 || p.warps * p.stages * (QS_TILE_BYTES + 8) > QS_MAX_SMEM)
    p = {16, 2, 1};
#endif
  return p;
}

static bool bits_ok(int bits) { return bits == 2 || bits == 4 || bits == 8 || bits == 16; }

template <int BITS>
static int quantize_bits(const float* x, int64_t rows, int64_t cols, int64_t sr, int64_t sc, int64_t block,
                         float eps, void* codes, uint8_t* packed, float* scales, float* dequant,
                         cudaStream_t st) {
  const int64_t numel = rows * cols;
  const bool whole = (block == numel);
  const bool contiguous = (sc == 1 && sr == cols) || (rows == 1 && sc == 1);
  auto aligned32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
  const bool fast = contiguous && aligned32(x) && (numel % 16 == 0) &&
                    (codes == nullptr || aligned16(codes)) && (packed == nullptr || aligned16(packed)) &&
                    (dequant == nullptr || aligned16(dequant)) &&
                    (whole || block == 32 || block == 64 || block == 128 || block == 256);
  if (fast) {
    const int grid = grid_for((numel + 1023) / 1024, 8, 8);  // 8 warps per CTA, one 1024-element tile per warp-iteration
    const bool full = (numel % 1024 == 0);
    // large tensors: one resident CTA per SM fed by bulk asynchronous copies (quant_stream_kernel)
    // only bit-packed codes wanted (and a width whose packing is worth specialising): the lean instantiation
    const bool pack_only = (BITS == 2 || BITS == 4) && codes == nullptr && dequant == nullptr && packed != nullptr;
    const QuantStreamPolicy sp = quant_stream_policy(BITS, pack_only);
    const int64_t ntiles = numel >> 10;
    const bool stream_path = full && sp.warps > 0 && ntiles >= (int64_t)kNumSMs * sp.warps * 2;
#define CB_QF_STREAM(LPB, PO)                                                                                     \
  do {                                                                                                            \
    static PerDeviceOnce once;                                                                                    \
    const int smem = sp.warps * sp.stages * (QS_TILE_BYTES + 8);                                                  \
    CB_TRY(opt_in_dynamic_smem(quant_stream_kernel<BITS, LPB, PO>, QS_MAX_SMEM, once));                           \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = dim3((unsigned)kNumSMs);                                                                        \
    cfg.blockDim = dim3((unsigned)(sp.warps * 32));                                                               \
    cfg.dynamicSmemBytes = (size_t)smem;                                                                          \
    cfg.stream = st;                                                                                              \
    cudaLaunchAttribute attr[1];                                                                                  \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                              \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                       \
    cfg.attrs = attr;                                                                                             \
    cfg.numAttrs = sp.pdl ? 1 : 0;                                                                                \
    CB_CUDA(cudaLaunchKernelEx(&cfg, quant_stream_kernel<BITS, LPB, PO>, x, ntiles, sp.stages, eps, codes, packed, \
                               scales, dequant));                                                                 \
  } while (0)
#define CB_QF_PO(LPB, PO)                                                                                         \
  do {                                                                                                            \
    if (stream_path) {                                                                                            \
      if constexpr (kStreamEveryWidth || (BITS == 2 && (PO))) CB_QF_STREAM(LPB, PO);                              \
    } else if (full) quant_fast_kernel<BITS, LPB, true, PO><<<grid, 256, 0, st>>>(x, numel, eps, codes, packed, scales, dequant); \
    else quant_fast_kernel<BITS, LPB, false, PO><<<grid, 256, 0, st>>>(x, numel, eps, codes, packed, scales, dequant);          \
  } while (0)
#define CB_QF(LPB)                                                                                                \
  do {                                                                                                            \
    if (pack_only) CB_QF_PO(LPB, (BITS == 2 || BITS == 4));                                                       \
    else CB_QF_PO(LPB, false);                                                                                    \
  } while (0)
    if (whole) {
      CB_CUDA(cudaMemsetAsync(scales, 0, sizeof(float), st));
      absmax_fast_kernel<<<grid_for(numel / 8, 256 * 8, 4), 256, 0, st>>>(x, numel, scales);
      CB_CHECK_LAUNCH();
      CB_QF(0);
    } else if (block == 32) {
      CB_QF(4);
    } else if (block == 64) {
      CB_QF(8);
    } else if (block == 128) {
      CB_QF(16);
    } else {
      CB_QF(32);
    }
#undef CB_QF
#undef CB_QF_PO
#undef CB_QF_STREAM
    CB_CHECK_LAUNCH();
    return CB_OK;
  }
  CB_CUDA(cudaMemsetAsync(scales, 0, sizeof(float) * (size_t)(numel / block), st));
  absmax_generic_kernel<<<grid_for(numel, 256, 8), 256, 0, st>>>(x, rows, cols, sr, sc, block, scales);
  CB_CHECK_LAUNCH();
  constexpr int G = BITS == 2 ? 4 : (BITS == 4 ? 2 : 1);
  quant_generic_kernel<BITS><<<grid_for((numel + G - 1) / G, 256, 8), 256, 0, st>>>(
      x, rows, cols, sr, sc, block, eps, codes, packed, scales, dequant);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

__global__ void clamp_scale_kernel(float* s, float eps) { s[0] = fmaxf(s[0], eps); }

}  // namespace cb

using namespace cb;

extern "C" size_t cb_packed_bytes(int64_t numel, int bits) {
  if (numel < 0) return 0;
  switch (bits) {
    case 2: return (size_t)((numel + 3) / 4);
    case 4: return (size_t)((numel + 1) / 2);
    case 8: return (size_t)numel;
    case 16: return (size_t)numel * 2;
    default: return 0;
  }
}

extern "C" int cb_quantize_f32(const float* x, int64_t rows, int64_t cols, int64_t stride_r, int64_t stride_c,
                               int bits, int64_t block, float eps, void* codes, uint8_t* packed,
                               float* scales, float* dequant, void* stream) {
  if (!bits_ok(bits)) return CB_ERR_BITS;
  if (rows < 0 || cols < 0) return CB_ERR_ARG;
  const int64_t numel = rows * cols;
  if (numel == 0) return CB_OK;
  if (x == nullptr || scales == nullptr) return CB_ERR_ARG;
  if (block == 0) block = numel;
  if (block < 0 || numel % block != 0) return CB_ERR_BLOCK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch (bits) {
    case 2: rc = quantize_bits<2>(x, rows, cols, stride_r, stride_c, block, eps, codes, packed, scales, dequant, st); break;
    case 4: rc = quantize_bits<4>(x, rows, cols, stride_r, stride_c, block, eps, codes, packed, scales, dequant, st); break;
    case 8: rc = quantize_bits<8>(x, rows, cols, stride_r, stride_c, block, eps, codes, packed, scales, dequant, st); break;
    default: rc = quantize_bits<16>(x, rows, cols, stride_r, stride_c, block, eps, codes, packed, scales, dequant, st); break;
  }
  if (rc != CB_OK) return rc;
  if (block == numel) {
    // whole-tensor mode: publish the clamped scale max(absmax, eps)
    clamp_scale_kernel<<<1, 1, 0, st>>>(scales, eps);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

template <int BITS>
static int dequant_bits(const void* codes, const uint8_t* packed, const float* scales, int64_t numel,
                        int64_t block, float* out, cudaStream_t st) {
  const bool fast = (BITS <= 8) && (numel % 16 == 0) && (block % 16 == 0) && aligned16(out) &&
                    (packed != nullptr ? aligned16(packed) : aligned16(codes));
  if (fast) {
    const int grid = grid_for(numel / 16, 256, 8);
    if (packed != nullptr) dequant_fast_kernel<BITS, true><<<grid, 256, 0, st>>>(codes, packed, scales, numel, block, out);
    else dequant_fast_kernel<BITS, false><<<grid, 256, 0, st>>>(codes, packed, scales, numel, block, out);
  } else {
    const int grid = grid_for(numel, 256, 8);
    if (packed != nullptr) dequant_generic_kernel<BITS, true><<<grid, 256, 0, st>>>(codes, packed, scales, numel, block, out);
    else dequant_generic_kernel<BITS, false><<<grid, 256, 0, st>>>(codes, packed, scales, numel, block, out);
  }
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_dequantize_f32(const void* codes, const uint8_t* packed, const float* scales,
                                 int64_t numel, int bits, int64_t block, float* out, void* stream) {
  if (!bits_ok(bits)) return CB_ERR_BITS;
  if ((codes == nullptr) == (packed == nullptr) || scales == nullptr || out == nullptr || numel < 0) return CB_ERR_ARG;
  if (numel == 0) return CB_OK;
  if (block == 0) block = numel;
  if (block < 0 || numel % block != 0) return CB_ERR_BLOCK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bits) {
    case 2: return dequant_bits<2>(codes, packed, scales, numel, block, out, st);
    case 4: return dequant_bits<4>(codes, packed, scales, numel, block, out, st);
    case 8: return dequant_bits<8>(codes, packed, scales, numel, block, out, st);
    default: return dequant_bits<16>(codes, packed, scales, numel, block, out, st);
  }
}

extern "C" int cb_pack_codes(const void* codes, int64_t numel, int bits, uint8_t* packed, void* stream) {
  if (!bits_ok(bits)) return CB_ERR_BITS;
  if (codes == nullptr || packed == nullptr || numel < 0) return CB_ERR_ARG;
  if (numel == 0) return CB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(numel, 256, 8);
  switch (bits) {
    case 2: pack_kernel<2><<<grid, 256, 0, st>>>(codes, numel, packed); break;
    case 4: pack_kernel<4><<<grid, 256, 0, st>>>(codes, numel, packed); break;
    case 8: pack_kernel<8><<<grid, 256, 0, st>>>(codes, numel, packed); break;
    default: pack_kernel<16><<<grid, 256, 0, st>>>(codes, numel, packed); break;
  }
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// ---------------------------------------------------------------- NormalFloat codebooks (nf4 / nf2)
// quantization.py:56-90: per block s = max(absmax, eps); index = #{thresholds t : x / s > t} with the
// thresholds midway between neighbouring levels; dequantised value = levels[index] * s.
namespace cb {
struct NfTable { float v[16]; int count; };

__global__ void __launch_bounds__(256)
nf_absmax_kernel(const float* __restrict__ x, int64_t numel, int64_t block, int chunks_per_block,
                 float* __restrict__ scales) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x / chunks_per_block;
  const int c = blockIdx.x % chunks_per_block;
  const int64_t per = (block + chunks_per_block - 1) / chunks_per_block;
  const int64_t lo = b * block + c * per, hi = min(b * block + block, lo + per);
  float a = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) a = fmaxf(a, fabsf(x[i]));
  a = block_max(a, red);
  if (threadIdx.x == 0) atomic_max_nonneg(scales + b, a);
}
__global__ void __launch_bounds__(256)
nf_code_kernel(const float* __restrict__ x, int64_t numel, int64_t block, float eps, const NfTable thr,
               float* __restrict__ scales, uint8_t* __restrict__ idx, int finalize_scales) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const int64_t b = i / block;
    const float s = fmaxf(scales[b], eps);
    const float w = __fdiv_rn(x[i], s);
    int k = 0;
#pragma unroll
    for (int t = 0; t < 15; ++t) k += (t < thr.count && w > thr.v[t]) ? 1 : 0;
    idx[i] = (uint8_t)k;
  }
  (void)finalize_scales;
}
__global__ void __launch_bounds__(256) nf_clamp_scales_kernel(float* __restrict__ scales, int64_t nblk, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblk) scales[i] = fmaxf(scales[i], eps);
}
__global__ void __launch_bounds__(256)
nf_dequant_kernel(const uint8_t* __restrict__ idx, const float* __restrict__ scales, int64_t numel, int64_t block,
                  const NfTable levels, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const int k = min((int)idx[i], levels.count - 1);
    out[i] = levels.v[k] * scales[i / block];
  }
}
}  // namespace cb

extern "C" int cb_quantize_nf_f32(const float* x, int64_t numel, int64_t block, const float* thresholds_host,
                                  int n_thresholds, float eps, uint8_t* idx, float* scales, void* stream) {
  using namespace cb;
  if (x == nullptr || idx == nullptr || scales == nullptr || thresholds_host == nullptr) return CB_ERR_ARG;
  if (numel <= 0 || block <= 0 || numel % block != 0) return CB_ERR_BLOCK;
  if (n_thresholds < 1 || n_thresholds > 15) return CB_ERR_BITS;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblk = numel / block;
  NfTable thr;
  thr.count = n_thresholds;
  for (int t = 0; t < 16; ++t) thr.v[t] = t < n_thresholds ? thresholds_host[t] : 0.f;
  CB_CUDA(cudaMemsetAsync(scales, 0, sizeof(float) * nblk, st));
  // enough CTAs per quantisation block to fill the machine when there are few (large) blocks
  int chunks = 1;
  if (nblk < 4 * kNumSMs) {
    const int64_t want = (4 * kNumSMs + nblk - 1) / nblk, most = (block + 1023) / 1024;
    chunks = (int)(want < most ? want : most);
  }
  if (chunks < 1) chunks = 1;
  if (nblk * chunks > (int64_t)INT32_MAX) return CB_ERR_ARG;
  nf_absmax_kernel<<<(unsigned)(nblk * chunks), 256, 0, st>>>(x, numel, block, chunks, scales);
  CB_CHECK_LAUNCH();
  nf_code_kernel<<<grid_for(numel, 256, 8), 256, 0, st>>>(x, numel, block, eps, thr, scales, idx, 0);
  CB_CHECK_LAUNCH();
  nf_clamp_scales_kernel<<<(unsigned)((nblk + 255) / 256), 256, 0, st>>>(scales, nblk, eps);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_dequantize_nf_f32(const uint8_t* idx, const float* scales, int64_t numel, int64_t block,
                                    const float* levels_host, int n_levels, float* out, void* stream) {
  using namespace cb;
  if (idx == nullptr || scales == nullptr || levels_host == nullptr || out == nullptr) return CB_ERR_ARG;
  if (numel <= 0 || block <= 0 || numel % block != 0) return CB_ERR_BLOCK;
  if (n_levels < 2 || n_levels > 16) return CB_ERR_BITS;
  NfTable lv;
  lv.count = n_levels;
  for (int t = 0; t < 16; ++t) lv.v[t] = t < n_levels ? levels_host[t] : 0.f;
  nf_dequant_kernel<<<grid_for(numel, 256, 8), 256, 0, (cudaStream_t)stream>>>(idx, scales, numel, block, lv, out);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" int cb_unpack_codes(const uint8_t* packed, int64_t numel, int bits, void* codes, void* stream) {
  if (!bits_ok(bits)) return CB_ERR_BITS;
  if (codes == nullptr || packed == nullptr || numel < 0) return CB_ERR_ARG;
  if (numel == 0) return CB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(numel, 256, 8);
  switch (bits) {
    case 2: unpack_kernel<2><<<grid, 256, 0, st>>>(packed, numel, codes); break;
    case 4: unpack_kernel<4><<<grid, 256, 0, st>>>(packed, numel, codes); break;
    case 8: unpack_kernel<8><<<grid, 256, 0, st>>>(packed, numel, codes); break;
    default: unpack_kernel<16><<<grid, 256, 0, st>>>(packed, numel, codes); break;
  }
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = alpha * A[M,K] * B[N,K]^T
//
// bf16 operands, both K-major (row-major with K contiguous), fp32 accumulation in tensor
// memory.  This is the contraction engine of the rank-r step (sketch, Gram and projection
// products of the randomized subspace iteration) and of the LPLR normal equations.
//
//   * one CTA per 128 x BN output tile (BN = 64/128/256) and K slice, 320 threads:
//       warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled 64-wide K blocks, one or two
//                   K blocks per instruction through a 2-D / 3-D tensor map)
//       warp 1      TMEM allocator + MMA issuer (one thread issues tcgen05.mma, K = 16 per
//                   instruction; tcgen05.commit releases the smem stage / signals the epilogue)
//       warps 2-9   epilogue: tcgen05.ld of the fp32 accumulator (lane = row; two groups of four warps,
//                   half of the columns each), optional column/row scaling, then fp32 stores, bf16 /
//                   transposed-bf16 outputs staged through shared memory and written as whole rows, or
//                   the raw partial tile of a K slice (summed in slice order by splitk_reduce_kernel:
//                   no floating-point atomics anywhere)
//   * STAGES-deep mbarrier ring between TMA and MMA; out-of-bounds rows/K are zero-filled by
//     TMA, so any M, N, K works as long as the leading dimensions are multiples of 8 elements.
//   * every mbarrier wait is bounded: a broken pipeline sets *error_flag and drains instead
//     of hanging the GPU.
//   * packed_linear_kernel (below) is the same pipeline with the B tile expanded from bit-packed codes
//     by the worker warps: the consumer of the packed decomposition.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"

namespace cb {

struct GemmTcArgs {
  int M, N, K;
  int kb_per_split;       // 64-wide K blocks per grid.z slice
  float alpha;
  float* C; int64_t ldc;             // fp32 out, may be null
  __nv_bfloat16* Cb; int64_t ldcb;   // bf16 out, row-major, may be null
  __nv_bfloat16* Ct; int64_t ldct;   // bf16 out, transposed (N x M), may be null
  const float* colscale;             // per output column, may be null
  const float* rowscale;             // per output row, may be null
  int staged;                        // bf16-only outputs whose alignment allows the shared-memory staged epilogue
  int loads_only;                    // measurement aid: run the TMA pipeline but issue no MMA (output undefined)
  long long* timing;                 // measurement aid: 8 clock64 stamps of CTA (0,0,0), or null
  int splits;                        // > 1: split-K, raw partial tiles go to `ws` for splitk_reduce_kernel
  float* ws;                         // [tile][split][BN/32][128][32] fp32 partial tiles
  int* error_flag;
};

constexpr int TC_BM = 128, TC_BK = 64;
// warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..9: epilogue.  A warp may only read the
// TMEM lanes of its quadrant (warp % 4), so the eight epilogue warps are two groups of four: group 0 drains
// the lower half of the accumulator's columns, group 1 the upper half (one warp per scheduler is latency
// bound at ~900 cycles per 32-column chunk; two halve the drain time of the 128 x 256 tile).
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_EPI_THREADS = 32 * TC_EPI_WARPS;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;

// One pipeline stage holds KB consecutive 64-wide K blocks of the A and B tiles, each block in the
// canonical 128-byte-swizzled K-major layout (rows 128 bytes apart), fetched by one TMA instruction
// per operand: issuing a tensor copy costs ~230 cycles of the producer thread whatever its size
// (measured, scripts/probe_gemm_phases.py), so narrow tiles are only fed fast enough with KB = 2.
template <int BN, int STAGES, int KB>
struct TcSmem {
  static constexpr int A_BLK = TC_BM * TC_BK * 2, B_BLK = BN * TC_BK * 2;
  static constexpr int A_BYTES = KB * A_BLK;
  static constexpr int B_BYTES = KB * B_BLK;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

// One thread's 32 consecutive output columns of one row: scaling, then the fp32 / bf16 /
// transposed-bf16 stores the caller asked for.
__device__ __forceinline__ void store_chunk(const GemmTcArgs& args, int row, int col0, float rs, float (&v)[32]) {
  if (row >= args.M || col0 >= args.N) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = v[j] * args.alpha * rs;
    if (args.colscale != nullptr && col0 + j < args.N) x *= args.colscale[col0 + j];
    v[j] = x;
  }
  const bool full = (col0 + 32 <= args.N);
  if (args.C != nullptr) {
    float* p = args.C + (int64_t)row * args.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (col0 + j < args.N) p[j] = v[j];
    }
  }
  if (args.Cb != nullptr) {
    __nv_bfloat16* p = args.Cb + (int64_t)row * args.ldcb + col0;
    if (full && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 w;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
        w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(p + j) = w;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (col0 + j < args.N) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (args.Ct != nullptr) {
    // transposed: lanes hold consecutive rows -> coalesced 64-byte runs per column
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < args.N) args.Ct[(int64_t)(col0 + j) * args.ldct + row] = __float2bfloat16_rn(v[j]);
  }
}

template <int BN, int STAGES, int KB>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcArgs args) {
  using S = TcSmem<BN, STAGES, KB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + STAGES * S::A_BYTES;
  const uint32_t bar_base = base + S::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  long long* stamps = (args.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? args.timing : nullptr;
  if (stamps != nullptr && threadIdx.x == 0) stamps[0] = clock64();               // kernel entry
  const int total_kb = (args.K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * args.kb_per_split;
  const int nkb = min(total_kb, kb0 + args.kb_per_split) - kb0;
  if (nkb <= 0) return;  // uniform per CTA
  const int nst = (nkb + KB - 1) / KB;   // pipeline iterations; the last may hold fewer than KB valid blocks

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  bool ok = true;

  if (warp == 0) {
    if (lane == 0) {
      if (stamps != nullptr) stamps[1] = clock64();                                // prologue done
      for (int i = 0; i < nst; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        if (!mbar_wait(empty_bar(s), ph ^ 1u)) { ok = false; break; }
        if (stamps != nullptr && i == nst - 1) stamps[2] = clock64();              // last load about to be issued
        mbar_expect_tx(full_bar(s), (uint32_t)(S::A_BYTES + S::B_BYTES));
        const int kb = kb0 + i * KB;
        if constexpr (KB == 1) {
          tma_load_2d(a_base + s * S::A_BYTES, &tmA, full_bar(s), kb * TC_BK, m_blk * TC_BM);
          tma_load_2d(b_base + s * S::B_BYTES, &tmB, full_bar(s), kb * TC_BK, n_blk * BN);
        } else {
          // blocks past the end of K are zero-filled; blocks past this CTA's K slice are loaded but not used
          tma_load_3d(a_base + s * S::A_BYTES, &tmA, full_bar(s), 0, m_blk * TC_BM, kb);
          tma_load_3d(b_base + s * S::B_BYTES, &tmB, full_bar(s), 0, n_blk * BN, kb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
      for (int i = 0; i < nst; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        if (!mbar_wait(full_bar(s), ph)) { ok = false; break; }
        if (stamps != nullptr && i == 0) stamps[3] = clock64();                    // first stage landed
        if (stamps != nullptr && i == nst - 1) stamps[4] = clock64();              // last stage landed
        tc_fence_after();
        if (args.loads_only) {
          if (i == 0) umma_bf16(tmem_base, make_smem_desc_sw128(a_base), make_smem_desc_sw128(b_base), idesc, 0u);
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar(s)) : "memory");
          continue;
        }
        const int valid = min(KB, nkb - i * KB);
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          if (b < valid) {
            const uint64_t adesc = make_smem_desc_sw128(a_base + s * S::A_BYTES + b * S::A_BLK);
            const uint64_t bdesc = make_smem_desc_sw128(b_base + s * S::B_BYTES + b * S::B_BLK);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr>>4) field
              umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                        (i > 0 || b > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
    const int qd = warp & 3;
    const int egrp = (warp - 2) >> 2;                                   // 0 or 1
    constexpr int EPI_COLS = BN / (TC_EPI_WARPS / 4);                   // columns drained by each group
    const int c_begin = egrp * EPI_COLS, c_end = c_begin + EPI_COLS;
    if (!mbar_wait(tmem_full_bar, 0)) ok = false;
    if (stamps != nullptr && threadIdx.x == 64) stamps[5] = clock64();             // accumulator complete
    ok = __all_sync(0xffffffffu, ok);
    tc_fence_after();
    const int trow = qd * 32 + lane;               // row inside the tile
    const int row = m_blk * TC_BM + trow;
    const float rs = (args.rowscale != nullptr && row < args.M) ? args.rowscale[row] : 1.f;
    if (args.splits <= 1 && args.staged) {
      // bf16-only outputs (the sketch / apply contractions): a thread owns a row of the accumulator, so
      // direct stores are 16-byte pieces 2*ld bytes apart (row-major) or 64-byte runs (transposed).  Stage
      // the tile through the now idle pipeline buffers in both orientations and write whole rows instead.
      constexpr int RM_STRIDE = BN * 2 + 16;        // bytes per staged row (row-major tile), padded against bank conflicts
      constexpr int T_STRIDE = TC_BM * 2 + 16;      // bytes per staged column (transposed tile)
      uint8_t* stage_rm = smem_raw + (base - smem_u32(smem_raw));
      uint8_t* stage_t = stage_rm + TC_BM * RM_STRIDE;
      if (ok) {
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r);
          const int col0 = n_blk * BN + c0;
          __align__(16) __nv_bfloat16 b[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]) * args.alpha * rs;
            if (args.colscale != nullptr && col0 + j < args.N) x *= args.colscale[col0 + j];
            b[j] = __float2bfloat16_rn(x);
          }
          if (args.Cb != nullptr) {
            uint4* dst = reinterpret_cast<uint4*>(stage_rm + trow * RM_STRIDE + c0 * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = *reinterpret_cast<const uint4*>(&b[8 * j]);
          }
          if (args.Ct != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              *reinterpret_cast<__nv_bfloat16*>(stage_t + (c0 + j) * T_STRIDE + trow * 2) = b[j];
          }
        }
      }
      if (stamps != nullptr && threadIdx.x == 64) stamps[8] = clock64();           // staged: tile in shared memory
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");     // the epilogue warps
      if (stamps != nullptr && threadIdx.x == 64) stamps[9] = clock64();           // staged: barrier passed
      if (ok) {
        const int et = threadIdx.x - 64;                 // 0 .. TC_EPI_THREADS - 1
        if (args.Cb != nullptr) {
          constexpr int UPR = BN / 8;                    // 16-byte units per row
#pragma unroll 4
          for (int u = et; u < TC_BM * UPR; u += TC_EPI_THREADS) {
            const int r_ = u / UPR, cu = u - r_ * UPR;
            const int grow = m_blk * TC_BM + r_, gcol = n_blk * BN + cu * 8;
            if (grow < args.M && gcol < args.N)
              *reinterpret_cast<uint4*>(args.Cb + (int64_t)grow * args.ldcb + gcol) =
                  *reinterpret_cast<const uint4*>(stage_rm + r_ * RM_STRIDE + cu * 16);
          }
        }
        if (args.Ct != nullptr) {
          constexpr int UPC = TC_BM / 4;                 // 8-byte units (4 rows) per column
#pragma unroll 4
          for (int u = et; u < BN * UPC; u += TC_EPI_THREADS) {
            const int c_ = u / UPC, ru = u - c_ * UPC;
            const int gcol = n_blk * BN + c_, grow = m_blk * TC_BM + ru * 4;
            if (gcol < args.N && grow < args.M)
              *reinterpret_cast<uint2*>(args.Ct + (int64_t)gcol * args.ldct + grow) =
                  *reinterpret_cast<const uint2*>(stage_t + c_ * T_STRIDE + ru * 8);
          }
        }
      }
    } else if (args.splits <= 1) {
      if (ok) {
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          store_chunk(args, row, n_blk * BN + c0, rs, v);
        }
      }
    } else {
      // Split-K without atomics, so that results do not depend on arrival order: every slice
      // parks its raw partial tile in the workspace ([tile][slice][BN/32][128][32] floats) and
      // splitk_reduce_kernel, launched right behind on the same stream, adds the slices up in
      // slice order and writes the outputs.
      const int tile = m_blk * gridDim.x + n_blk;
      if (ok) {
        float* mine = args.ws + ((size_t)tile * args.splits + blockIdx.z) * (BN * TC_BM) + (size_t)trow * 32;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r);
          float4* dst = reinterpret_cast<float4*>(mine + (size_t)(c0 / 32) * (TC_BM * 32));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(dst + j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
      }
    }
  }
  if (!ok && args.error_flag != nullptr) atomicExch(args.error_flag, 1);
  if (stamps != nullptr && threadIdx.x == 64) stamps[6] = clock64();               // epilogue stores issued
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
  if (stamps != nullptr && threadIdx.x == 0) stamps[7] = clock64();               // exit
}

// ---------------------------------------------------------------- packed-code consumer
// y[T, m] = gs * ( (s / lv) * x[T, n] * C[m, n]^T  +  t[T, r] * L[m, r]^T )      
// C are the b-bit codes of Q read straight from their packed form (never materialised in HBM), t = x R^T
// comes from a preceding contraction.  Same pipeline as gemm_tc_kernel, except that the B tile of a code
// K block is produced by the worker warps: thread = tile row, 64 codes expanded through a shared-memory
// lookup table into the bf16 integers -lv .. lv (exact) and written in the 128-byte-swizzled K-major layout
// the MMA descriptor expects; the whole-tensor scale is applied once in the epilogue.  The rank-r part runs
// as extra K blocks (A = t, B = L, both by TMA) into a second TMEM accumulator.
struct PackedLinearArgs {
  int T, m, n, r;
  int nkb_codes, nkb_lr;
  const uint8_t* packed;      // m x n codes, BITS each, MSB-first
  const float* q_scale;       // device scalar
  float gs, lv;
  float* y; int64_t ldy;
  int* error_flag;
  long long* timing;          // measurement aid (cb_set_gemm_timing): 8 clock64 stamps of CTA (0,0), or null
};

constexpr int PL_BN = 128, PL_STAGES = 3, PL_KB = 2, PL_WORKERS = 256;
constexpr int PL_THREADS = 64 + PL_WORKERS;
// One stage = PL_KB (2) consecutive 64-wide K blocks of the A and B tiles: the expansion warps pay one
// proxy fence + barrier round trip per 128 codes instead of per 64 (that round trip, not the expansion
// itself, bounded the first version at ~1000 cycles per K block).
struct PlSmem {
  static constexpr int A_BLK = TC_BM * TC_BK * 2, B_BLK = PL_BN * TC_BK * 2;
  static constexpr int A_BYTES = PL_KB * A_BLK, B_BYTES = PL_KB * B_BLK;
  static constexpr int BAR_OFF = PL_STAGES * (A_BYTES + B_BYTES);
  static constexpr int LUT_OFF = BAR_OFF + (2 * PL_STAGES + 1) * 8 + 16;
  static constexpr int TOTAL = LUT_OFF + 2048 + 1024;
};

template <int BITS>
__global__ void __launch_bounds__(PL_THREADS, 1)
packed_linear_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmT,
                     const __grid_constant__ CUtensorMap tmL, const PackedLinearArgs args) {
  using S = PlSmem;
  constexpr int ROW_BYTES = TC_BK * BITS / 8;   // packed bytes of one tile row per K block (16 / 32 / 64)
  constexpr int LVI = (1 << (BITS - 1)) - 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base, b_base = base + PL_STAGES * S::A_BYTES;
  const uint32_t bar_base = base + S::BAR_OFF;
  auto full_bar = [&](int s_) { return bar_base + 8u * s_; };
  auto empty_bar = [&](int s_) { return bar_base + 8u * (PL_STAGES + s_); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * PL_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * PL_STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + S::BAR_OFF + 8 * (2 * PL_STAGES + 1));
  uint32_t* lut = reinterpret_cast<uint32_t*>(base_ptr + S::LUT_OFF);   // [16][32]: value x lane

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;      // tile of m (output columns), tile of T (rows)
  const int ns_codes = (args.nkb_codes + PL_KB - 1) / PL_KB, ns_lr = (args.nkb_lr + PL_KB - 1) / PL_KB;
  const int ns = ns_codes + ns_lr;                       // pipeline iterations (stages of PL_KB K blocks)
  long long* stamps = (args.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0) ? args.timing : nullptr;
  if (stamps != nullptr && threadIdx.x == 0) stamps[0] = clock64();              // entry

  // Expansion table, one private copy per lane so that a lookup never conflicts (entry (v, lane) lives in bank
  // `lane`): a byte-indexed table with random indices cost ~5-way bank conflicts and made the kernel
  // shared-memory bound.  2-bit: nibble -> two codes as a bf16 pair; 4-bit: nibble -> one code.
  for (int e = threadIdx.x; e < 16 * 32; e += blockDim.x) {
    const int v = e >> 5;
    uint32_t entry;
    if (BITS == 2) {
      const uint32_t lo = __bfloat16_as_ushort(__int2bfloat16_rn(((v >> 2) & 3) - LVI));   // first code: high bits
      const uint32_t hi = __bfloat16_as_ushort(__int2bfloat16_rn((v & 3) - LVI));
      entry = lo | (hi << 16);
    } else {
      entry = __bfloat16_as_ushort(__int2bfloat16_rn(v - LVI));
    }
    lut[e] = entry;
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmL); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s_ = 0; s_ < PL_STAGES; ++s_) { mbar_init(full_bar(s_), 1 + PL_WORKERS / 32); mbar_init(empty_bar(s_), 1); }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * PL_BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  bool ok = true;

  if (warp == 0) {
    if (lane == 0) {
      if (stamps != nullptr) stamps[1] = clock64();                                // prologue done
      for (int i = 0; i < ns; ++i) {
        const int s_ = i % PL_STAGES;
        const uint32_t ph = (uint32_t)(i / PL_STAGES) & 1u;
        if (!mbar_wait(empty_bar(s_), ph ^ 1u)) { ok = false; break; }
        if (i < ns_codes) {
          // both K blocks of x with one 3-D copy (a block past n / 64 is zero-filled)
          mbar_expect_tx(full_bar(s_), (uint32_t)S::A_BYTES);
          tma_load_3d(a_base + s_ * S::A_BYTES, &tmX, full_bar(s_), 0, m_blk * TC_BM, i * PL_KB);
        } else {
          const int kb0 = (i - ns_codes) * PL_KB;
          const int valid = min(PL_KB, args.nkb_lr - kb0);
          mbar_expect_tx(full_bar(s_), (uint32_t)(valid * (S::A_BLK + S::B_BLK)));
          for (int b2 = 0; b2 < valid; ++b2) {
            tma_load_2d(a_base + s_ * S::A_BYTES + b2 * S::A_BLK, &tmT, full_bar(s_), (kb0 + b2) * TC_BK, m_blk * TC_BM);
            tma_load_2d(b_base + s_ * S::B_BYTES + b2 * S::B_BLK, &tmL, full_bar(s_), (kb0 + b2) * TC_BK, n_blk * PL_BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, PL_BN);
      for (int i = 0; i < ns; ++i) {
        const int s_ = i % PL_STAGES;
        const uint32_t ph = (uint32_t)(i / PL_STAGES) & 1u;
        if (!mbar_wait(full_bar(s_), ph)) { ok = false; break; }
        if (stamps != nullptr && i == 0) stamps[3] = clock64();                    // first stage complete
        if (stamps != nullptr && i == ns - 1) stamps[4] = clock64();               // last stage complete
        tc_fence_after();
        const bool lr = i >= ns_codes;
        const uint32_t acc = tmem_base + (lr ? (uint32_t)PL_BN : 0u);
        const int valid = lr ? min(PL_KB, args.nkb_lr - (i - ns_codes) * PL_KB) : min(PL_KB, args.nkb_codes - i * PL_KB);
        const bool first_stage = lr ? (i == ns_codes) : (i == 0);
#pragma unroll
        for (int b2 = 0; b2 < PL_KB; ++b2) {
          if (b2 < valid) {
            const uint64_t adesc = make_smem_desc_sw128(a_base + s_ * S::A_BYTES + b2 * S::A_BLK);
            const uint64_t bdesc = make_smem_desc_sw128(b_base + s_ * S::B_BYTES + b2 * S::B_BLK);
#pragma unroll
            for (int k2 = 0; k2 < TC_BK / 16; ++k2)
              umma_bf16(acc, adesc + (uint64_t)(2 * k2), bdesc + (uint64_t)(2 * k2), idesc,
                        (!first_stage || b2 > 0 || k2 > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s_));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- worker warps: expand the packed codes of this CTA's 128 rows of C, stage by stage; two threads
    //      per row, thread `sub` expanding K block 2 * stage + sub (64 codes = ROW_BYTES contiguous bytes)
    const int wt = threadIdx.x - 64;                      // 0 .. 255
    const int t = wt >> 1, sub = wt & 1;                  // tile row, K block inside the stage
    const int64_t crow = (int64_t)n_blk * PL_BN + t;      // row of C (= output column of y)
    const bool row_ok = crow < args.m;
    const uint8_t* src = args.packed + (row_ok ? crow : 0) * ((int64_t)args.n * BITS / 8);
    constexpr int NW = ROW_BYTES / 16;                    // 128-bit words per thread and stage (1 / 2 / 4)
    // the packed codes of the next PF stages stay in registers: one L2 round trip is several stages long
    constexpr int PF = BITS == 2 ? 4 : (BITS == 4 ? 3 : 2);
    uint4 buf[PF][NW];
    auto fetch = [&](int stage, uint4 (&dst)[NW]) {
      const int kb = stage * PL_KB + sub;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        dst[w] = make_uint4(0, 0, 0, 0);
        if (row_ok && kb < args.nkb_codes) dst[w] = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)kb * ROW_BYTES) + w);
      }
    };
#pragma unroll
    for (int j = 0; j < PF; ++j) fetch(j, buf[j]);
    for (int i = 0; i < ns; ++i) {
      const int s_ = i % PL_STAGES;
      const uint32_t ph = (uint32_t)(i / PL_STAGES) & 1u;
      uint4 cur[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) cur[w] = buf[0][w];
#pragma unroll
      for (int j = 0; j + 1 < PF; ++j)
#pragma unroll
        for (int w = 0; w < NW; ++w) buf[j][w] = buf[j + 1][w];
      fetch(i + PF, buf[PF - 1]);
      if (!mbar_wait(empty_bar(s_), ph ^ 1u)) { ok = false; break; }
      if (i < ns_codes && i * PL_KB + sub < args.nkb_codes) {
        uint8_t* dst = base_ptr + (b_base - base) + s_ * S::B_BYTES + sub * S::B_BLK + t * 128;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(cur);
#pragma unroll
        for (int c = 0; c < 8; ++c) {                      // chunk: 8 consecutive K elements = BITS packed bytes
          uint32_t out[4];
          if (BITS == 2) {
            // 2 packed bytes = 4 nibbles = 4 output words (two codes each)
            const uint32_t hw = (words[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;     // bytes 2c (low), 2c + 1 (high)
            out[0] = lut[((hw >> 4) & 15u) * 32 + lane];
            out[1] = lut[(hw & 15u) * 32 + lane];
            out[2] = lut[((hw >> 12) & 15u) * 32 + lane];
            out[3] = lut[((hw >> 8) & 15u) * 32 + lane];
          } else if (BITS == 4) {
            const uint32_t w4 = words[c];                                          // bytes 4c .. 4c + 3
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
              const uint32_t byte = (w4 >> (8 * b2)) & 255u;
              out[b2] = lut[(byte >> 4) * 32 + lane] | (lut[(byte & 15u) * 32 + lane] << 16);
            }
          } else {
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
              const uint32_t w8 = words[2 * c + (b2 >> 1)];
              const int c0_ = (int)((w8 >> (16 * (b2 & 1))) & 255u) - LVI, c1_ = (int)((w8 >> (16 * (b2 & 1) + 8)) & 255u) - LVI;
              out[b2] = (uint32_t)__bfloat16_as_ushort(__int2bfloat16_rn(c0_)) |
                        ((uint32_t)__bfloat16_as_ushort(__int2bfloat16_rn(c1_)) << 16);
            }
          }
          *reinterpret_cast<uint4*>(dst + ((c ^ (t & 7)) << 4)) = make_uint4(out[0], out[1], out[2], out[3]);
        }
        fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(s_)) : "memory");
    }
    // ---- epilogue
    if (stamps != nullptr && threadIdx.x == 64) stamps[2] = clock64();             // expansion loop done
    const int qd = warp & 3;
    if (!mbar_wait(tmem_full_bar, 0)) ok = false;
    if (stamps != nullptr && threadIdx.x == 64) stamps[5] = clock64();             // accumulators complete
    ok = __all_sync(0xffffffffu, ok);
    tc_fence_after();
    const int row = m_blk * TC_BM + qd * 32 + lane;
    const float sc = args.q_scale[0] / args.lv;
    const int egrp = (warp - 2) >> 2;                     // two groups of four warps, half of the columns each
    if (ok) {
#pragma unroll 1
      for (int c0 = egrp * (PL_BN / 2); c0 < (egrp + 1) * (PL_BN / 2); c0 += 32) {
        uint32_t a1[32], a2[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, a1);
        if (args.nkb_lr > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(PL_BN + c0), a2);
        const int col0 = n_blk * PL_BN + c0;
        if (row < args.T && col0 < args.m) {
          float* p = args.y + (int64_t)row * args.ldy + col0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = sc * __uint_as_float(a1[j]);
            if (args.nkb_lr > 0) x += __uint_as_float(a2[j]);
            v[j] = args.gs * x;
          }
          if (col0 + 32 <= args.m && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < args.m) p[j] = v[j];
          }
        }
      }
    }
  }
  if (!ok && args.error_flag != nullptr) atomicExch(args.error_flag, 1);
  if (stamps != nullptr && threadIdx.x == 64) stamps[6] = clock64();               // epilogue stores issued
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * PL_BN);
  }
  if (stamps != nullptr && threadIdx.x == 0) stamps[7] = clock64();               // exit
}

// Measurement aid (bench / scripts only): back-to-back tcgen05.mma on one resident shared-memory
// stage, no TMA and no per-stage barriers, to read off what the tensor pipe itself sustains for a
// 128 x N x 16 instruction with both operands in shared memory.  out[0] = cycles for `n_mma` MMAs.
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
mma_rate_probe_kernel(int n_mma, int distinct_k, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 4 * TC_BM * TC_BK * 2;
  const uint32_t bar = b_base + 4 * BN * TC_BK * 2;
  const uint32_t tmem_slot = bar + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: zeros are as good as anything for timing
  for (uint32_t i = threadIdx.x; i < (uint32_t)(4 * (TC_BM + BN) * TC_BK * 2 / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int st = (i / 4) % 4, k = i % 4;          // walk 4 stages x 4 K steps like the real main loop
      const int sk = distinct_k ? k : 0, ss = distinct_k ? st : 0;
      const uint64_t adesc = make_smem_desc_sw128(a_base + ss * (TC_BM * TC_BK * 2)) + (uint64_t)(2 * sk);
      const uint64_t bdesc = make_smem_desc_sw128(b_base + ss * (BN * TC_BK * 2)) + (uint64_t)(2 * sk);
      umma_bf16(tmem_base, adesc, bdesc, idesc, i > 0 ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BN); }
}

// Second half of a split-K contraction: one thread per 4 output columns sums the slices of its
// partial tile in slice order, then scales and stores like the unsplit epilogue.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const GemmTcArgs args, int bn, int tiles_n, int64_t total4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  // idx enumerates float4 slots in workspace order: [tile][chunk][row in tile][8]
  const int j4 = (int)(idx & 7);
  const int trow = (int)((idx >> 3) & (TC_BM - 1));
  const int64_t rest = idx >> 10;                       // tile * (bn/32) + chunk
  const int chunks = bn / 32;
  const int chunk = (int)(rest % chunks);
  const int tile = (int)(rest / chunks);
  const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
  const size_t tile_elems = (size_t)bn * TC_BM;
  const float* src = args.ws + (size_t)tile * args.splits * tile_elems + (size_t)chunk * (TC_BM * 32) +
                     (size_t)trow * 32 + (size_t)j4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int z = 0; z < args.splits; ++z) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * tile_elems));
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  const int row = m_blk * TC_BM + trow;
  const int col = n_blk * bn + chunk * 32 + j4 * 4;
  if (row >= args.M || col >= args.N) return;
  const float rs = args.alpha * (args.rowscale != nullptr ? args.rowscale[row] : 1.f);
  float v[4] = {acc.x * rs, acc.y * rs, acc.z * rs, acc.w * rs};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (args.colscale != nullptr && col + j < args.N) v[j] *= args.colscale[col + j];
  if (args.C != nullptr) {
    float* p = args.C + (int64_t)row * args.ldc + col;
    if (col + 4 <= args.N && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else
      for (int j = 0; j < 4; ++j)
        if (col + j < args.N) p[j] = v[j];
  }
  for (int j = 0; j < 4; ++j) {
    if (col + j >= args.N) break;
    if (args.Cb != nullptr) args.Cb[(int64_t)row * args.ldcb + col + j] = __float2bfloat16_rn(v[j]);
    if (args.Ct != nullptr) args.Ct[(int64_t)(col + j) * args.ldct + row] = __float2bfloat16_rn(v[j]);
  }
}

// ---------------------------------------------------------------- host side
// rows x K bf16 matrix, K contiguous, leading dimension ld (elements); box = 64 (K) x box_rows
static int make_tmap_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return CB_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CB_OK : CB_ERR_ARG;
}

// The same matrix seen as (64 K elements) x rows x (K / 64 blocks), so that one box fetches `kblocks`
// consecutive 64-wide K blocks of `box_rows` rows; they land one after the other in shared memory,
// each in the same swizzled layout as a 2-D box.  K must be a multiple of 64 (a ragged last block
// would run into the next row instead of being zero-filled); blocks at or past K / 64 are zero-filled.
static int make_tmap_bf16_kblocks(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows,
                                  int kblocks) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return CB_ERR_UNSUPPORTED;
  if (K % TC_BK != 0) return CB_ERR_ARG;
  cuuint64_t gdim[3] = {(cuuint64_t)TC_BK, (cuuint64_t)rows, (cuuint64_t)(K / TC_BK)};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)TC_BK * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, (cuuint32_t)kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CB_OK : CB_ERR_ARG;
}

template <int BN, int STAGES, int KB>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& args, int splits, cudaStream_t st) {
  using S = TcSmem<BN, STAGES, KB>;
  static PerDeviceOnce once;
  CB_TRY(opt_in_dynamic_smem(gemm_tc_kernel<BN, STAGES, KB>, S::TOTAL, once));
  dim3 grid((unsigned)((args.N + BN - 1) / BN), (unsigned)((args.M + TC_BM - 1) / TC_BM), (unsigned)splits);
  gemm_tc_kernel<BN, STAGES, KB><<<grid, TC_THREADS, S::TOTAL, st>>>(ta, tb, args);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

thread_local ExecPolicy tl_policy;
#ifdef CB_MEASURE
long long* g_timing = nullptr;   // measurement aid, see cb_set_gemm_timing
int g_staged_epilogue = 1;       // cb_set_gemm_staged_epilogue(0) falls back to direct stores
int g_kblocks = 2;               // K blocks per TMA instruction for the narrow tiles (cb_set_gemm_kblocks: 1 or 2)
#else
constexpr long long* g_timing = nullptr;
constexpr int g_staged_epilogue = 1;
constexpr int g_kblocks = 2;
#endif

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb) {
  return M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31) && lda % 8 == 0 &&
         ldb % 8 == 0 && lda >= K && ldb >= K && aligned16(A) && aligned16(B);
}

// splitk <= 0: choose automatically so the grid fills the machine.  Split-K needs the scratch in
// `sw` (see SplitWs) and runs as two launches (partial tiles, then splitk_reduce_kernel); outputs are
// overwritten, never accumulated into, and the sum over K slices is taken in slice order, so a given
// (shape, grid policy) always produces the same bits.
int gemm_tc(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A, int64_t lda,
            const __nv_bfloat16* B, int64_t ldb, float* C, int64_t ldc, __nv_bfloat16* Cb, int64_t ldcb,
            __nv_bfloat16* Ct, int64_t ldct, const float* colscale, const float* rowscale, int splitk,
            int* error_flag, int* splits_used, cudaStream_t st, const SplitWs* sw, int probe_flags) {
  if (!gemm_tc_supported(M, N, K, A, lda, B, ldb)) return CB_ERR_UNSUPPORTED;
  if (debug_skip() & 4) return CB_OK;
  // largest N tile that still yields >= ~0.8 waves of CTAs; otherwise the smallest tile
  // (most CTAs) and, if the caller allows it, a K split on top
  const int total_kb = (int)((K + TC_BK - 1) / TC_BK);
  const int64_t mt = (M + TC_BM - 1) / TC_BM;
  // tl_policy.target_ctas: how many CTAs a grid should reach before a fatter N tile / fewer K splits are
  // preferred.  Every gemm_tc CTA owns an SM (192 KiB of shared memory), so grids from different
  // streams only overlap when they are small: ~32-CTA grids let four layers' contractions run side by
  // side and move fewer bytes per flop (throughput mode, several layers in flight); ~120-CTA grids
  // fill the machine for one layer (latency mode, the default).  Part of the call (cb_caldera_params.exec_mode).
  const int min_ctas = tl_policy.target_ctas;
  int bn = 64;
  if (N >= 192 && mt * ((N + 255) / 256) >= min_ctas) bn = 256;
  else if (N >= 96 && mt * ((N + 127) / 128) >= min_ctas) bn = 128;
  else if (N >= 192 && splitk == 1 && mt * ((N + 63) / 64) < 32) bn = 256;   // tiny grids: fewer, fatter CTAs
  const int64_t tiles = mt * ((N + bn - 1) / bn);
  int splits = splitk;
  if (splits <= 0) {
    splits = 1;
    const int fill = min_ctas >= 120 ? kNumSMs : min_ctas;
    if (tiles < fill) splits = (int)((fill + tiles - 1) / tiles);
    const int max_splits = total_kb / 4 > 0 ? total_kb / 4 : 1;  // at least 4 K blocks per slice
    if (splits > max_splits) splits = max_splits;
  }
  if (splits > total_kb) splits = total_kb;
  if (splits > 1) {
    // bounded by the scratch the caller provided
    const size_t per_split = (size_t)tiles * TC_BM * bn * sizeof(float);
    const int fit = (sw != nullptr && sw->buf != nullptr) ? (int)(sw->bytes / per_split) : 1;
    if (splits > fit) splits = fit;
  }
  if (splits < 1) splits = 1;
  int kb_per = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per - 1) / kb_per;
  if (splits_used != nullptr) *splits_used = splits;
  const bool loads_only = (probe_flags & 1) != 0;
  // two K blocks per TMA instruction where the tile is narrow enough for the stage to stay <= 64 KiB
  const int kbs = (bn <= 128 && K % TC_BK == 0 && g_kblocks >= 2) ? 2 : 1;
  if (kbs > 1 && kb_per % kbs != 0 && splits > 1) {
    kb_per = (kb_per + kbs - 1) / kbs * kbs;              // K slices start on a stage boundary
    splits = (total_kb + kb_per - 1) / kb_per;
    if (splits_used != nullptr) *splits_used = splits;
  }
  CUtensorMap ta, tb;
  if (kbs > 1) {
    CB_TRY(make_tmap_bf16_kblocks(&ta, A, M, K, lda, TC_BM, kbs));
    CB_TRY(make_tmap_bf16_kblocks(&tb, B, N, K, ldb, bn, kbs));
  } else {
    CB_TRY(make_tmap_bf16(&ta, A, M, K, lda, TC_BM));
    CB_TRY(make_tmap_bf16(&tb, B, N, K, ldb, bn));
  }
  GemmTcArgs args;
  args.M = (int)M; args.N = (int)N; args.K = (int)K; args.kb_per_split = kb_per; args.alpha = alpha;
  args.C = C; args.ldc = ldc; args.Cb = Cb; args.ldcb = ldcb; args.Ct = Ct; args.ldct = ldct;
  args.colscale = colscale; args.rowscale = rowscale; args.error_flag = error_flag;
  args.splits = splits; args.ws = splits > 1 ? sw->buf : nullptr;
  args.loads_only = loads_only ? 1 : 0;
  // staged epilogue: bf16 outputs only, whole 8-column / 4-row units inside the matrix, vector-aligned
  args.staged = (g_staged_epilogue != 0 && splits == 1 && C == nullptr && (Cb != nullptr || Ct != nullptr) &&
                 (Cb == nullptr || (N % 8 == 0 && ldcb % 8 == 0 && aligned16(Cb))) &&
                 (Ct == nullptr || (M % 4 == 0 && ldct % 4 == 0 && (reinterpret_cast<uintptr_t>(Ct) & 7u) == 0))) ? 1 : 0;
  args.timing = g_timing;
  if (bn == 256) CB_TRY((launch_tc<256, 4, 1>(ta, tb, args, splits, st)));
  else if (bn == 128 && kbs == 2) CB_TRY((launch_tc<128, 3, 2>(ta, tb, args, splits, st)));
  else if (bn == 128) CB_TRY((launch_tc<128, 6, 1>(ta, tb, args, splits, st)));
  else if (kbs == 2) CB_TRY((launch_tc<64, 4, 2>(ta, tb, args, splits, st)));
  else CB_TRY((launch_tc<64, 8, 1>(ta, tb, args, splits, st)));
  if (splits > 1) {
    const int64_t total4 = tiles * TC_BM * (bn / 4);
    splitk_reduce_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(args, bn, (int)((N + bn - 1) / bn), total4);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

// fp32 -> bf16 (optionally transposed and/or scaled) conversions around the tensor-core path
__global__ void __launch_bounds__(256)
to_bf16_kernel(const float* __restrict__ X_, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* __restrict__ Y_,
               int64_t ldy, const float* __restrict__ colscale_, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.y;
  const float* __restrict__ X = boff(X_, bo);
  __nv_bfloat16* __restrict__ Y = boff(Y_, bo);
  const float* __restrict__ colscale = boff(colscale_, bo);
  const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    float v = X[r * ldx + c];
    if (colscale != nullptr) v *= colscale[c];
    Y[r * ldy + c] = __float2bfloat16_rn(v);
  }
}
// Yt (cols x rows, bf16) = X^T, tiled through shared memory
__global__ void __launch_bounds__(256)
to_bf16_t_kernel(const float* __restrict__ X_, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* __restrict__ Yt_,
                 int64_t ldyt, const float* __restrict__ colscale_, int64_t bstride) {
  const int64_t bo = bstride * blockIdx.z;
  const float* __restrict__ X = boff(X_, bo);
  __nv_bfloat16* __restrict__ Yt = boff(Yt_, bo);
  const float* __restrict__ colscale = boff(colscale_, bo);
  __shared__ float tile[32][33];
  const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = by + k, c = bx + tx;
    float v = 0.f;
    if (r < rows && c < cols) { v = X[r * ldx + c]; if (colscale != nullptr) v *= colscale[c]; }
    tile[k][tx] = v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t c = bx + k, r = by + tx;
    if (r < rows && c < cols) Yt[c * ldyt + r] = __float2bfloat16_rn(tile[tx][k]);
  }
}

int to_bf16(const float* X, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* Y, int64_t ldy,
            __nv_bfloat16* Yt, int64_t ldyt, const float* colscale, cudaStream_t st, const Bt& bt) {
  if (Y != nullptr) {
    dim3 grid1((unsigned)grid_for(rows * cols, 256 * 4, bt.n > 1 ? 2 : 8), (unsigned)bt.n);
    to_bf16_kernel<<<grid1, 256, 0, st>>>(X, rows, cols, ldx, Y, ldy, colscale, bt.stride);
    CB_CHECK_LAUNCH();
  }
  if (Yt != nullptr) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)bt.n);
    to_bf16_t_kernel<<<grid, 256, 0, st>>>(X, rows, cols, ldx, Yt, ldyt, colscale, bt.stride);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

}  // namespace cb

// C ABI: exported so the tensor-core path can be validated in isolation against a reference GEMM
extern "C" size_t cb_gemm_bf16_tn_workspace_bytes(void) { return cb::kSplitWsBytes; }

extern "C" int cb_gemm_bf16_tn(int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                               const void* B_bf16, int64_t ldb, float* C, int64_t ldc, int splitk, int probe_flags,
                               int* error_flag, void* workspace, size_t workspace_bytes, void* stream) {
  if (A_bf16 == nullptr || B_bf16 == nullptr || C == nullptr) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int splits = 0;
  cb::SplitWs sw;
  if (workspace != nullptr && workspace_bytes > 0) {
    sw.buf = reinterpret_cast<float*>(workspace); sw.bytes = workspace_bytes;
  } else if (splitk > 1) {
    return CB_ERR_WORKSPACE;
  }
  return cb::gemm_tc(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(A_bf16), lda,
                     reinterpret_cast<const __nv_bfloat16*>(B_bf16), ldb, C, ldc, nullptr, 0, nullptr, 0, nullptr,
                     nullptr, splitk, error_flag, &splits, st, &sw, probe_flags);
}

// ---- packed-code consumer: y = x (Q + L R)^T * global_scale from the packed decomposition
namespace cb {
struct PackedLinearPlan {
  __nv_bfloat16 *xb, *Rb, *Lb, *tb;
  SplitWs sw;
  size_t bytes;
};
static PackedLinearPlan plan_packed_linear(uint8_t* base, int64_t T, int64_t m, int64_t n, int64_t r) {
  PackedLinearPlan P{};
  size_t off = 0;
  auto take = [&](size_t b) { void* p = base != nullptr ? base + off : nullptr; off += (b + 255) / 256 * 256; return p; };
  P.xb = reinterpret_cast<__nv_bfloat16*>(take(sizeof(__nv_bfloat16) * T * n));
  if (r > 0) {
    P.Rb = reinterpret_cast<__nv_bfloat16*>(take(sizeof(__nv_bfloat16) * r * n));
    P.Lb = reinterpret_cast<__nv_bfloat16*>(take(sizeof(__nv_bfloat16) * m * r));
    P.tb = reinterpret_cast<__nv_bfloat16*>(take(sizeof(__nv_bfloat16) * T * r));
    P.sw.buf = reinterpret_cast<float*>(take(kSplitWsBytes));
    P.sw.bytes = kSplitWsBytes;
  }
  P.bytes = off;
  return P;
}
}  // namespace cb

extern "C" size_t cb_packed_linear_workspace_bytes(int64_t T, int64_t m, int64_t n, int64_t r) {
  if (T <= 0 || m <= 0 || n <= 0 || r < 0) return 0;
  return cb::plan_packed_linear(nullptr, T, m, n, r).bytes + 256;
}

extern "C" int cb_packed_linear_f32(const float* x, int64_t T, int64_t n, const uint8_t* q_packed, int q_bits,
                                    const float* q_scale, const float* L, const float* R, int64_t m, int64_t r,
                                    float global_scale, float* y, int* error_flag, void* ws, size_t ws_bytes,
                                    void* stream) {
  using namespace cb;
  if (x == nullptr || q_packed == nullptr || q_scale == nullptr || y == nullptr || ws == nullptr) return CB_ERR_ARG;
  if (T <= 0 || m <= 0 || n <= 0 || r < 0 || (r > 0 && (L == nullptr || R == nullptr))) return CB_ERR_ARG;
  if (q_bits != 2 && q_bits != 4 && q_bits != 8) return CB_ERR_BITS;
  if (n % 64 != 0 || r % 8 != 0 || !aligned16(q_packed) || T >= (1ll << 31) || m >= (1ll << 31)) return CB_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  PackedLinearPlan P = plan_packed_linear(base, T, m, n, r);
  if ((size_t)(base - reinterpret_cast<uint8_t*>(ws)) + P.bytes > ws_bytes) return CB_ERR_WORKSPACE;
  CB_TRY(to_bf16(x, T, n, n, P.xb, n, nullptr, 0, nullptr, st));
  if (r > 0) {
    CB_TRY(to_bf16(R, r, n, n, P.Rb, n, nullptr, 0, nullptr, st));
    CB_TRY(to_bf16(L, m, r, r, P.Lb, r, nullptr, 0, nullptr, st));
    // t[T, r] = x R^T (bf16 out, consumed as the A operand of the rank-r K blocks)
    CB_TRY(gemm_tc(T, r, n, 1.f, P.xb, n, P.Rb, n, nullptr, 0, P.tb, r, nullptr, 0, nullptr, nullptr, 0, error_flag,
                   nullptr, st, &P.sw));
  }
  CUtensorMap tx, tt, tl;
  CB_TRY(make_tmap_bf16_kblocks(&tx, P.xb, T, n, n, TC_BM, PL_KB));
  if (r > 0) {
    CB_TRY(make_tmap_bf16(&tt, P.tb, T, r, r, TC_BM));
    CB_TRY(make_tmap_bf16(&tl, P.Lb, m, r, r, PL_BN));
  } else {
    CB_TRY(make_tmap_bf16(&tt, P.xb, T, n, n, TC_BM));     // unused placeholders (no rank-r stages)
    tl = tt;
  }
  PackedLinearArgs a;
  a.T = (int)T; a.m = (int)m; a.n = (int)n; a.r = (int)r;
  a.nkb_codes = (int)(n / TC_BK); a.nkb_lr = (int)((r + TC_BK - 1) / TC_BK);
  a.packed = q_packed; a.q_scale = q_scale; a.gs = global_scale; a.lv = (float)((1 << (q_bits - 1)) - 1);
  a.y = y; a.ldy = m; a.error_flag = error_flag; a.timing = g_timing;
  dim3 grid((unsigned)((m + PL_BN - 1) / PL_BN), (unsigned)((T + TC_BM - 1) / TC_BM));
#define CB_PL(BITS)                                                                                                  \
  do {                                                                                                               \
    static PerDeviceOnce once;                                                                                       \
    CB_TRY(opt_in_dynamic_smem(packed_linear_kernel<BITS>, PlSmem::TOTAL, once));                                    \
    packed_linear_kernel<BITS><<<grid, PL_THREADS, PlSmem::TOTAL, st>>>(tx, tt, tl, a);                              \
  } while (0)
  if (q_bits == 2) CB_PL(2);
  else if (q_bits == 4) CB_PL(4);
  else CB_PL(8);
#undef CB_PL
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// Same contraction with the bf16 epilogues the layer driver uses: Cb (M x N, row-major) and/or Ct (N x M,
// transposed), optional per-column / per-row scaling.  Exported so that the staged epilogue can be tested.
extern "C" int cb_gemm_bf16_tn_bf16out(int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                                       const void* B_bf16, int64_t ldb, void* Cb_bf16, int64_t ldcb, void* Ct_bf16,
                                       int64_t ldct, const float* colscale, const float* rowscale, int exec_mode,
                                       int* error_flag, void* stream) {
  if (A_bf16 == nullptr || B_bf16 == nullptr || (Cb_bf16 == nullptr && Ct_bf16 == nullptr)) return CB_ERR_ARG;
  cb::PolicyScope policy(exec_mode);
  return cb::gemm_tc(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(A_bf16), lda,
                     reinterpret_cast<const __nv_bfloat16*>(B_bf16), ldb, nullptr, 0,
                     reinterpret_cast<__nv_bfloat16*>(Cb_bf16), ldcb, reinterpret_cast<__nv_bfloat16*>(Ct_bf16), ldct,
                     colscale, rowscale, 1, error_flag, nullptr, (cudaStream_t)stream, nullptr, 0);
}

#ifdef CB_MEASURE
extern "C" void cb_set_gemm_staged_epilogue(int on) { cb::g_staged_epilogue = on != 0 ? 1 : 0; }
#endif

extern "C" int cb_convert_bf16(const float* X, int64_t rows, int64_t cols, int64_t ldx, void* Y_bf16, int64_t ldy,
                               void* Yt_bf16, int64_t ldyt, const float* colscale, void* stream) {
  if (X == nullptr || rows <= 0 || cols <= 0) return CB_ERR_ARG;
  return cb::to_bf16(X, rows, cols, ldx, reinterpret_cast<__nv_bfloat16*>(Y_bf16), ldy,
                     reinterpret_cast<__nv_bfloat16*>(Yt_bf16), ldyt, colscale, (cudaStream_t)stream);
}

#ifdef CB_MEASURE
// Measurement aid: cycles for n_mma back-to-back 128 x bn x 16 bf16 MMAs on every SM of a `grid`-CTA launch
// (out: 2 device int64: [0] issue + drain, [1] issue only).
extern "C" int cb_probe_mma_rate(int bn, int n_mma, int distinct_k, int grid, void* out_cycles, void* stream) {
  if (out_cycles == nullptr || n_mma < 1 || grid < 1) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 4 * (cb::TC_BM + bn) * cb::TC_BK * 2 + 64 + 1024;
#define CB_PROBE(BN)                                                                                              \
  do {                                                                                                            \
    CB_CUDA(cudaFuncSetAttribute(cb::mma_rate_probe_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    cb::mma_rate_probe_kernel<BN><<<grid, cb::TC_THREADS, smem, st>>>(n_mma, distinct_k, (long long*)out_cycles);  \
  } while (0)
  if (bn == 64) CB_PROBE(64);
  else if (bn == 128) CB_PROBE(128);
  else if (bn == 256) CB_PROBE(256);
  else return CB_ERR_ARG;
#undef CB_PROBE
  CB_CHECK_LAUNCH();
  return CB_OK;
}

extern "C" void cb_set_gemm_timing(void* stamps_dev) { cb::g_timing = reinterpret_cast<long long*>(stamps_dev); }
extern "C" void cb_set_gemm_kblocks(int n) { cb::g_kblocks = n == 1 ? 1 : 2; }
#endif  // CB_MEASURE

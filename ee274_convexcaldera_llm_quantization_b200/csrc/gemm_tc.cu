// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] (+)= alpha * A[M,K] * B[N,K]^T
//
// bf16 operands, both K-major (row-major with K contiguous), fp32 accumulation in tensor
// memory.  This is the contraction engine of the rank-r step (sketch, Gram and projection
// products of the randomized subspace iteration) and of the LPLR normal equations.
//
//   * one CTA per 128 x BN output tile (BN = 64/128/256) and K-split, 192 threads:
//       warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled 64-wide K blocks)
//       warp 1      TMEM allocator + MMA issuer (one thread issues tcgen05.mma, K = 16 per
//                   instruction; tcgen05.commit releases the smem stage / signals the epilogue)
//       warps 2-5   epilogue: tcgen05.ld of the fp32 accumulator (lane = row), optional
//                   column/row scaling, fp32 / bf16 / transposed-bf16 stores or split-K atomics
//   * STAGES-deep mbarrier ring between TMA and MMA; out-of-bounds rows/K are zero-filled by
//     TMA, so any M, N, K works as long as the leading dimensions are multiples of 8 elements.
//   * every mbarrier wait is bounded: a broken pipeline sets *error_flag and drains instead
//     of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "internal.h"

namespace cb {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: returns false after ~2 s so a protocol bug cannot hang the device
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) return false;
  }
  return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t gets row (lane base + t), register j = column j
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp: start>>4 [0,14), LBO>>4 [16,30),
//  SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format=F32 [4,6), a/b_format=BF16 [7,10)/[10,13), K-major both,
// n_dim=N>>3 [17,23), m_dim=M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GemmTcArgs {
  int M, N, K;
  int kb_per_split;       // 64-wide K blocks per grid.z slice
  float alpha;
  float* C; int64_t ldc;             // fp32 out (store or atomic add), may be null
  __nv_bfloat16* Cb; int64_t ldcb;   // bf16 out, row-major, may be null (ignored when atomic)
  __nv_bfloat16* Ct; int64_t ldct;   // bf16 out, transposed (N x M), may be null (ignored when atomic)
  const float* colscale;             // per output column, may be null
  const float* rowscale;             // per output row, may be null
  int atomic;                        // 1: split-K partial sums via red.global.add
  int* error_flag;
};

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_THREADS = 192;

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcArgs args) {
  using S = TcSmem<BN, STAGES>;
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
  const int total_kb = (args.K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * args.kb_per_split;
  const int nkb = min(total_kb, kb0 + args.kb_per_split) - kb0;
  if (nkb <= 0) return;  // uniform per CTA

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
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        if (!mbar_wait(empty_bar(s), ph ^ 1u)) { ok = false; break; }
        mbar_expect_tx(full_bar(s), (uint32_t)(S::A_BYTES + S::B_BYTES));
        tma_load_2d(a_base + s * S::A_BYTES, &tmA, full_bar(s), (kb0 + i) * TC_BK, m_blk * TC_BM);
        tma_load_2d(b_base + s * S::B_BYTES, &tmB, full_bar(s), (kb0 + i) * TC_BK, n_blk * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        if (!mbar_wait(full_bar(s), ph)) { ok = false; break; }
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(a_base + s * S::A_BYTES);
        const uint64_t bdesc = make_smem_desc_sw128(b_base + s * S::B_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr>>4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
    const int qd = warp & 3;
    if (!mbar_wait(tmem_full_bar, 0)) ok = false;
    ok = __all_sync(0xffffffffu, ok);
    tc_fence_after();
    const int row = m_blk * TC_BM + qd * 32 + lane;
    const float rs = (args.rowscale != nullptr && row < args.M) ? args.rowscale[row] : 1.f;
    if (ok) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r);
        const int col0 = n_blk * BN + c0;
        if (row < args.M && col0 < args.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]) * args.alpha * rs;
            if (args.colscale != nullptr && col0 + j < args.N) x *= args.colscale[col0 + j];
            v[j] = x;
          }
          const bool full = (col0 + 32 <= args.N);
          if (args.atomic) {
            float* p = args.C + (int64_t)row * args.ldc + col0;
            if (full && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < args.N) atomicAdd(p + j, v[j]);
            }
          } else {
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
        }
      }
    }
  }
  if (!ok && args.error_flag != nullptr) atomicExch(args.error_flag, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

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

template <int BN, int STAGES>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& args, int splits, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  static bool attr = false;
  if (!attr) {
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr = true;
  }
  dim3 grid((unsigned)((args.N + BN - 1) / BN), (unsigned)((args.M + TC_BM - 1) / TC_BM), (unsigned)splits);
  gemm_tc_kernel<BN, STAGES><<<grid, TC_THREADS, S::TOTAL, st>>>(ta, tb, args);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

int g_target_ctas = -1;

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb) {
  return M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31) && lda % 8 == 0 &&
         ldb % 8 == 0 && lda >= K && ldb >= K && aligned16(A) && aligned16(B);
}

// splitk <= 0: choose automatically so the grid fills the machine.  With split-K the fp32
// output C must be zero-initialised by the caller (partials are atomically added) unless
// `accumulate` semantics are wanted.
int gemm_tc(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A, int64_t lda,
            const __nv_bfloat16* B, int64_t ldb, float* C, int64_t ldc, __nv_bfloat16* Cb, int64_t ldcb,
            __nv_bfloat16* Ct, int64_t ldct, const float* colscale, const float* rowscale, int splitk,
            int* error_flag, int* splits_used, cudaStream_t st) {
  if (!gemm_tc_supported(M, N, K, A, lda, B, ldb)) return CB_ERR_UNSUPPORTED;
  // largest N tile that still yields >= ~0.8 waves of CTAs; otherwise the smallest tile
  // (most CTAs) and, if the caller allows it, a K split on top
  const int total_kb = (int)((K + TC_BK - 1) / TC_BK);
  const int64_t mt = (M + TC_BM - 1) / TC_BM;
  // g_target_ctas: how many CTAs a grid should reach before a fatter N tile / fewer K splits are
  // preferred.  Every gemm_tc CTA owns an SM (192 KiB of shared memory), so grids from different
  // streams only overlap when they are small: ~32-CTA grids let four layers' contractions run side by
  // side and move fewer bytes per flop (throughput mode, several layers in flight); ~120-CTA grids
  // fill the machine for one layer (latency mode, the default).  Set by cb_set_gemm_target_ctas or
  // the CB_GEMM_MIN_CTAS environment variable.
  if (g_target_ctas < 0) {
    const char* e = getenv("CB_GEMM_MIN_CTAS");
    g_target_ctas = (e != nullptr && atoi(e) > 0) ? atoi(e) : 120;
  }
  const int min_ctas = g_target_ctas;
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
  if (splits < 1) splits = 1;
  int kb_per = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per - 1) / kb_per;
  if (splits > 1 && C == nullptr) return CB_ERR_ARG;
  if (splits_used != nullptr) *splits_used = splits;
  CUtensorMap ta, tb;
  CB_TRY(make_tmap_bf16(&ta, A, M, K, lda, TC_BM));
  CB_TRY(make_tmap_bf16(&tb, B, N, K, ldb, bn));
  GemmTcArgs args;
  args.M = (int)M; args.N = (int)N; args.K = (int)K; args.kb_per_split = kb_per; args.alpha = alpha;
  args.C = C; args.ldc = ldc; args.Cb = Cb; args.ldcb = ldcb; args.Ct = Ct; args.ldct = ldct;
  args.colscale = colscale; args.rowscale = rowscale; args.atomic = splits > 1 ? 1 : 0; args.error_flag = error_flag;
  if (bn == 256) return launch_tc<256, 4>(ta, tb, args, splits, st);
  if (bn == 128) return launch_tc<128, 6>(ta, tb, args, splits, st);
  return launch_tc<64, 8>(ta, tb, args, splits, st);
}

// fp32 -> bf16 (optionally transposed and/or scaled) conversions around the tensor-core path
__global__ void __launch_bounds__(256)
to_bf16_kernel(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* __restrict__ Y,
               int64_t ldy, const float* __restrict__ colscale) {
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
to_bf16_t_kernel(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* __restrict__ Yt,
                 int64_t ldyt, const float* __restrict__ colscale) {
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
            __nv_bfloat16* Yt, int64_t ldyt, const float* colscale, cudaStream_t st) {
  if (Y != nullptr) {
    to_bf16_kernel<<<grid_for(rows * cols, 256 * 4, 8), 256, 0, st>>>(X, rows, cols, ldx, Y, ldy, colscale);
    CB_CHECK_LAUNCH();
  }
  if (Yt != nullptr) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    to_bf16_t_kernel<<<grid, 256, 0, st>>>(X, rows, cols, ldx, Yt, ldyt, colscale);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

}  // namespace cb

// C ABI: exported so the tensor-core path can be validated in isolation against a reference GEMM
extern "C" int cb_gemm_bf16_tn(int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16, int64_t lda,
                               const void* B_bf16, int64_t ldb, float* C, int64_t ldc, int splitk, int* error_flag,
                               void* stream) {
  if (A_bf16 == nullptr || B_bf16 == nullptr || C == nullptr) return CB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int splits = 0;
  // split-K accumulates atomically: clear the output first (row by row when ldc > N)
  if (splitk != 1) CB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  return cb::gemm_tc(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(A_bf16), lda,
                     reinterpret_cast<const __nv_bfloat16*>(B_bf16), ldb, C, ldc, nullptr, 0, nullptr, 0, nullptr,
                     nullptr, splitk, error_flag, &splits, st);
}

extern "C" int cb_convert_bf16(const float* X, int64_t rows, int64_t cols, int64_t ldx, void* Y_bf16, int64_t ldy,
                               void* Yt_bf16, int64_t ldyt, const float* colscale, void* stream) {
  if (X == nullptr || rows <= 0 || cols <= 0) return CB_ERR_ARG;
  return cb::to_bf16(X, rows, cols, ldx, reinterpret_cast<__nv_bfloat16*>(Y_bf16), ldy,
                     reinterpret_cast<__nv_bfloat16*>(Yt_bf16), ldyt, colscale, (cudaStream_t)stream);
}

extern "C" void cb_set_gemm_target_ctas(int n) { cb::g_target_ctas = n > 0 ? n : 120; }

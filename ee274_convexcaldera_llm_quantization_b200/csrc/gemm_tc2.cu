// Batched, persistent tcgen05 GEMM on CTA pairs (sm_100a):
//     C[b][M, N] = alpha * A[b][M, K] * B[b][N, K]^T        b = 0 .. batch-1
//
// bf16 operands, both K-major, fp32 accumulation in tensor memory.  This is the contraction engine of the rank-r step
// when many same-shape layers advance in lock step (driver.cu, batched layer driver): one launch covers the same
// contraction of every layer in the batch, so a launch has hundreds of tiles even though one layer's sketch
// contraction has only 16.
//
//   * A cluster of two CTAs (one TPC) owns a 256 x N_TILE output tile: `tcgen05.mma.cta_group::2`, issued by one thread
//     of the leader CTA, M = 256 (each CTA holds 128 rows of A and of the accumulator), N = N_TILE <= 256 chosen per
//     launch as a multiple of 16 (the sketch contractions have N = q = 224: no padded columns are multiplied).  Each
//     CTA loads only half of the B tile; the tensor cores read the other half from the peer's shared memory.  Per
//     flop this halves the B-operand traffic from L2 compared with the 128 x 256 single-CTA tile of gemm_tc.cu, which
//     is what bounds a skinny contraction (every m-tile re-reads all of B) once all SMs are busy.
//   * Persistent with a dynamic tile scheduler: warp 3 of the leader CTA hands out tiles (batch-major, m fastest) -- the
//     first one per cluster statically, the following ones from a global atomic counter -- through a 4-deep ring in
//     the shared memory of both CTAs (remote store + release.cluster arrive), so a cluster that starts late (because
//     its SMs were still busy with another stream's Cholesky or element-wise kernel) simply takes fewer tiles instead
//     of delaying the launch by its whole static share.  Without a counter the assignment is the fixed stride.  The shared-
//     memory ring (STAGES x {A 16 KiB, B <= 16 KiB}) runs across tile boundaries and tensor memory holds TWO
//     accumulators (2 x 256 columns), so the epilogue of tile i (tcgen05.ld -> registers -> global) overlaps the main
//     loop of tile i + 1.
//   * Warp roles (384 threads): warp 0 TMA producer for A, warp 2 TMA producer for B (issuing a tensor copy costs the
//     issuing thread ~230 cycles whatever its size, so the two operands are issued from two warps), warp 1 TMEM
//     allocator + MMA issuer, warps 4-11 epilogue (a warp may only read the TMEM lanes of its quadrant, warp % 4; two
//     groups of four split the columns).
//   * Barriers: full[s] lives in the leader CTA and counts the bytes of both CTAs' copies (the copies of the second CTA
//     signal the leader's barrier: `cp.async.bulk.tensor...cta_group::2` with the peer bit of the barrier address
//     cleared); empty[s] and tmem_full[a] are signalled in both CTAs by multicast `tcgen05.commit`; tmem_empty[a] lives
//     in the leader and collects one arrival per epilogue warp of both CTAs (remote `mbarrier.arrive` through `mapa`).
//   * The batch is a third tensor-map dimension (stride = distance between two layers' operands), so one pair of
//     tensor maps serves the whole launch.  Out-of-bounds rows / K are zero-filled by TMA.
//   * Every mbarrier wait is bounded: a broken pipeline sets *error_flag and drains instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"

namespace cb {

constexpr int G2_BM = 128;                 // rows per CTA (256 per cluster)
constexpr int G2_BK = 64;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;          // 16 KiB
constexpr int G2_B_BYTES = 128 * G2_BK * 2;            // 16 KiB reserved per stage (N_TILE / 2 <= 128 rows used)
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 32 * (4 + G2_EPI_WARPS);    // 384
constexpr int G2_SQ = 4;                                   // depth of the tile-id ring
constexpr int G2_STAGES = 6;
constexpr int G2_BAR_OFF = G2_STAGES * (G2_A_BYTES + G2_B_BYTES);
constexpr int G2_NBAR = 2 * G2_STAGES + 4 + 2 * G2_SQ;     // full/empty ring, tmem full/empty x2, sched full/empty ring
constexpr int G2_EPI_OFF = G2_BAR_OFF + 1024;              // per epilogue warp: a 32 x 32 fp32 box (4 KiB, 128-byte swizzle) for TMA stores
constexpr int G2_EPI_BYTES = 32 * 32 * 4;
constexpr int G2_SMEM = G2_EPI_OFF + G2_EPI_WARPS * G2_EPI_BYTES + 1024;
static_assert(G2_NBAR * 8 + 16 + 4 * G2_SQ <= 1024, "barrier block overlaps the epilogue staging buffers");
static_assert(G2_SMEM <= 227 * 1024, "shared memory budget");

struct Gemm2Args {
  int M, N, K, batch;
  int tiles_m, tiles_n;       // 256-row tiles, N_TILE-column tiles per batch item
  int n_tile;                 // MMA N: multiple of 16, <= 256
  float alpha;
  // outputs (any subset); byte strides between batch items
  float* C; int64_t ldc; int64_t sC;
  __nv_bfloat16* Cb; int64_t ldcb; int64_t sCb;     // row-major bf16
  __nv_bfloat16* Ct; int64_t ldct; int64_t sCt;     // transposed (N x M) bf16
  // optional split-bf16 copy of C for a later K = 3 N' contraction ([hi | hi | lo] x [hi | lo | hi]^T ~ fp32 product):
  // split_mode 1: row i of Cs (M x 3N) = [hi(C[i,:]) | hi(C[i,:]) | lo(C[i,:])]
  // split_mode 2: row j of Cs (N x 3M) = [hi(C[:,j]) | lo(C[:,j]) | hi(C[:,j])]   (transposed)
  __nv_bfloat16* Cs; int64_t sCs; int split_mode;
  const float* colscale; int64_t sCol;              // per output column, may be null
  const float* rowscale; int64_t sRow;              // per output row, may be null
  int* error_flag; int64_t sFlag;
  int* tile_counter;             // 2 zeroed ints (next tile, clusters done; the kernel leaves them zero) or null = static
  int c_tma;                     // fp32 C is the only output and is TMA-storable: staged per warp in shared memory
};

// 32 consecutive output columns of one row: scaling, then the fp32 / bf16 / transposed-bf16 stores asked for
__device__ __forceinline__ void g2_store_chunk(const Gemm2Args& a, int b, int row, int col0, float rs, float (&v)[32]) {
  if (row >= a.M || col0 >= a.N) return;
  const float* cs = boff(a.colscale, a.sCol * b);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = v[j] * a.alpha * rs;
    if (cs != nullptr && col0 + j < a.N) x *= cs[col0 + j];
    v[j] = x;
  }
  const bool full = (col0 + 32 <= a.N);
  if (a.C != nullptr) {
    float* p = boff(a.C, a.sC * b) + (int64_t)row * a.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (col0 + j < a.N) p[j] = v[j];
    }
  }
  if (a.Cb != nullptr) {
    __nv_bfloat16* p = boff(a.Cb, a.sCb * b) + (int64_t)row * a.ldcb + col0;
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
        if (col0 + j < a.N) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (a.Ct != nullptr) {
    // transposed: the lanes of a warp hold consecutive rows -> 64-byte runs per output column
    __nv_bfloat16* p = boff(a.Ct, a.sCt * b);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) p[(int64_t)(col0 + j) * a.ldct + row] = __float2bfloat16_rn(v[j]);
  }
  if (a.Cs != nullptr) {
    __nv_bfloat16* p = boff(a.Cs, a.sCs * b);
    if (a.split_mode == 1) {
      __nv_bfloat16* q = p + (int64_t)row * 3 * a.N + col0;
      if (full && a.N % 8 == 0 && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0)) {
        // 32 consecutive columns of one row: three 64-byte pieces (hi, hi, lo), 128-bit stores
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[j + 2 * k]), h1 = __float2bfloat16_rn(v[j + 2 * k + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[j + 2 * k] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v[j + 2 * k + 1] - __bfloat162float(h1));
            h[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            l[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(q + j) = hv;
          *reinterpret_cast<uint4*>(q + a.N + j) = hv;
          *reinterpret_cast<uint4*>(q + 2 * a.N + j) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          if (col0 + j < a.N) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
            q[j] = hi; q[a.N + j] = hi; q[2 * a.N + j] = lo;
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < a.N) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
          const __nv_bfloat16 lo = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
          __nv_bfloat16* q = p + (int64_t)(col0 + j) * 3 * a.M + row;
          q[0] = hi; q[a.M] = lo; q[2 * a.M] = hi;
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const Gemm2Args args) {
  extern __shared__ uint8_t smem_raw[];
  // identical in both CTAs of the pair (same kernel, same dynamic shared memory window), which cta_group::2 relies on
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + G2_STAGES * G2_A_BYTES;
  const uint32_t bar_base = base + G2_BAR_OFF;
  const uint32_t epi_base = base + G2_EPI_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G2_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * G2_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * G2_STAGES + 2 + a); };
  auto sched_full_bar = [&](int q) { return bar_base + 8u * (2 * G2_STAGES + 4 + q); };
  auto sched_empty_bar = [&](int q) { return bar_base + 8u * (2 * G2_STAGES + 4 + G2_SQ + q); };
  const uint32_t tmem_slot = bar_base + 8u * G2_NBAR;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const uint32_t sched_tile = tmem_slot + 16u;             // G2_SQ ints: the tile ids handed out by the scheduler
  volatile int* sched_tile_ptr = reinterpret_cast<volatile int*>(smem_raw + (sched_tile - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tiles_per_item = args.tiles_m * args.tiles_n;
  const int total_tiles = args.batch * tiles_per_item;
  const int nkb = (args.K + G2_BK - 1) / G2_BK;
  const int b_rows = args.n_tile >> 1;                         // B rows held by each CTA
  const uint32_t stage_tx = 2u * (uint32_t)(G2_A_BYTES + b_rows * G2_BK * 2);   // bytes both CTAs deliver per stage

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); if (args.c_tma) tma_prefetch_desc(&tmC); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < G2_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 2 * G2_EPI_WARPS); }
      // consumers of a tile id: 2 producers + 8 epilogue warps per CTA, + the MMA issuer of the leader
      for (int q = 0; q < G2_SQ; ++q) { mbar_init(sched_full_bar(q), 1); mbar_init(sched_empty_bar(q), 2 * (2 + G2_EPI_WARPS) + 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, 512);
  }
  tc_fence_before();
  cluster_sync_all();            // barrier inits visible to the peer; also the CTA-wide sync for the TMEM address
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  bool ok = true;
  // every consumer role walks the same sequence of tile ids: wait for the i-th id, read it, release the ring slot
  auto next_tile = [&](int i, bool arrive) -> int {
    const int q = i % G2_SQ;
    const uint32_t ph = (uint32_t)(i / G2_SQ) & 1u;
    if (!mbar_wait_cluster(sched_full_bar(q), ph)) { ok = false; return -1; }
    const int t = sched_tile_ptr[q];
    if (arrive) mbar_arrive_release_cluster(sched_empty_bar(q), 0);
    return t;
  };
  auto release_tile = [&](int i) { mbar_arrive_release_cluster(sched_empty_bar(i % G2_SQ), 0); };

  if (warp == 3) {
    // ---- tile scheduler (leader CTA)
    if (leader && lane == 0) {
      int* counter = args.tile_counter;
      for (int i = 0;; ++i) {
        const int q = i % G2_SQ;
        const uint32_t ph = (uint32_t)(i / G2_SQ) & 1u;
        if (!mbar_wait_cluster(sched_empty_bar(q), ph ^ 1u)) ok = false;
        int t;
        if (i == 0) t = cluster_id;
        else if (counter != nullptr) t = atomicAdd(counter, 1) + n_clusters;
        else t = cluster_id + i * n_clusters;
        if (t >= total_tiles || !ok) t = -1;
        sched_tile_ptr[q] = t;
        st_shared_cluster_u32(sched_tile + 4u * q, 1, (uint32_t)t);
        mbar_arrive_release_cluster(sched_full_bar(q), 0);
        mbar_arrive_release_cluster(sched_full_bar(q), 1);
        if (t < 0) break;
      }
      if (counter != nullptr) {
        // the last cluster to run out of tiles leaves the counters zeroed for the next launch on this stream
        if (atomicAdd(counter + 1, 1) == n_clusters - 1) { counter[0] = 0; counter[1] = 0; __threadfence(); }
      }
    }
  } else if (warp == 0 || warp == 2) {
    // ---- TMA producers: warp 0 loads this CTA's 128 rows of A, warp 2 its half of the B tile
    if (lane == 0) {
      const bool is_a = warp == 0;
      uint32_t cnt = 0;
      for (int i = 0; ok; ++i) {
        const int t = next_tile(i, true);
        if (t < 0) break;
        const int b = t / tiles_per_item, rem = t - b * tiles_per_item;
        const int n_blk = rem / args.tiles_m, m_blk = rem - n_blk * args.tiles_m;
        const int row0 = is_a ? (m_blk * 2 * G2_BM + (int)cta_rank * G2_BM) : (n_blk * args.n_tile + (int)cta_rank * b_rows);
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int s = cnt % G2_STAGES;
          const uint32_t ph = (cnt / G2_STAGES) & 1u;
          if (!mbar_wait(empty_bar(s), ph ^ 1u)) { ok = false; break; }
          if (is_a) {
            if (leader) mbar_expect_tx(full_bar(s), stage_tx);
            tma_load_3d_2sm(a_base + s * G2_A_BYTES, &tmA, full_bar(s), kb * G2_BK, row0, b);
          } else {
            tma_load_3d_2sm(b_base + s * G2_B_BYTES, &tmB, full_bar(s), kb * G2_BK, row0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: one thread of the leader CTA drives the tensor cores of both SMs
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * G2_BM, args.n_tile);
      uint32_t cnt = 0;
      for (int it = 0; ok; ++it) {
        if (next_tile(it, true) < 0) break;
        const int acc = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        if (!mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u)) { ok = false; break; }   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int s = cnt % G2_STAGES;
          const uint32_t ph = (cnt / G2_STAGES) & 1u;
          if (!mbar_wait(full_bar(s), ph)) { ok = false; break; }
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(a_base + s * G2_A_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(b_base + s * G2_B_BYTES);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_bf16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(empty_bar(s));           // both CTAs' producers may refill the stage once these MMAs have read it
        }
        if (!ok) break;
        umma_commit_2sm(tmem_full_bar(acc));       // accumulator complete: wakes the epilogue warps of both CTAs
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31; two groups of four warps split the columns
    const int qd = warp & 3;
    const int egrp = (warp - 4) >> 2;                                  // 0 or 1
    const int n_chunks = (args.n_tile + 31) >> 5;
    const int c_split = (n_chunks + 1) >> 1;
    const int c_begin = egrp == 0 ? 0 : c_split, c_end = egrp == 0 ? c_split : n_chunks;
    for (int it = 0;; ++it) {
      const int t = next_tile(it, false);        // every lane reads the id ...
      __syncwarp();
      if (lane == 0) release_tile(it);           // ... before the warp gives the ring slot back
      if (t < 0) break;
      const int b = t / tiles_per_item, rem = t - b * tiles_per_item;
      const int n_blk = rem / args.tiles_m, m_blk = rem - n_blk * args.tiles_m;
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      if (ok && !mbar_wait(tmem_full_bar(acc), acc_ph)) ok = false;
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const int row = m_blk * 2 * G2_BM + (int)cta_rank * G2_BM + qd * 32 + lane;
      const float* rsp = boff(args.rowscale, args.sRow * b);
      const float rs = (rsp != nullptr && row < args.M) ? rsp[row] : 1.f;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * 256 + c * 32), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int col0 = n_blk * args.n_tile + c * 32;
        if (args.c_tma) {
          // fp32 output through a swizzled shared-memory box and one TMA store per 32 x 32 block: a direct store would
          // cost the load/store unit 32 line transactions per instruction (every lane owns a different row)
          const float* cs = boff(args.colscale, args.sCol * b);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j] * args.alpha * rs;
            if (cs != nullptr && col0 + j < args.N) x *= cs[col0 + j];
            v[j] = x;
          }
          const uint32_t buf = epi_base + (uint32_t)(warp - 4) * G2_EPI_BYTES;
          if (lane == 0) bulk_wait_group_read0();          // the previous store has read this buffer
          __syncwarp();
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            st_shared_v4(buf + (uint32_t)lane * 128u + (uint32_t)((q4 ^ (lane & 7)) << 4), v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2],
                         v[4 * q4 + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && col0 < args.N) {
            tma_store_3d(&tmC, buf, col0, m_blk * 2 * G2_BM + (int)cta_rank * G2_BM + qd * 32, b);
            bulk_commit_group();
          }
        } else {
          g2_store_chunk(args, b, row, col0, rs, v);
        }
      }
      // this warp has read its part of the accumulator: one arrival on the leader's tmem_empty barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(acc), 0);
    }
    if (args.c_tma && lane == 0) bulk_wait_group0();       // this warp's stores have landed
  }
  if (!ok && args.error_flag != nullptr) atomicExch(boff(args.error_flag, 0), 1);
  tc_fence_before();
  cluster_sync_all();            // both CTAs are done with the pair's shared state (peer smem, barriers, TMEM)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
// rows x K bf16 matrix per batch item, K contiguous, leading dimension ld (elements), `batch` items `sbytes` apart:
// 3-D map {K, rows, batch}, box {64, box_rows, 1}, 128-byte swizzle
static int make_tmap_bf16_batched(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int64_t batch,
                                  int64_t sbytes, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return CB_ERR_UNSUPPORTED;
  if (batch <= 1) { batch = 1; sbytes = ((rows * ld * 2 + 15) / 16) * 16; }
  if (sbytes % 16 != 0 || sbytes <= 0) return CB_ERR_ARG;
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)sbytes};
  cuuint32_t box[3] = {(cuuint32_t)G2_BK, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CB_OK : CB_ERR_ARG;
}

// fp32 output, rows x cols per batch item, leading dimension ld (elements): 3-D map {cols, rows, batch}, box {32, 32, 1},
// 128-byte swizzle (one box row = 32 floats = 128 bytes)
static int make_tmap_f32_out(CUtensorMap* map, float* ptr, int64_t rows, int64_t cols, int64_t ld, int64_t batch, int64_t sbytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return CB_ERR_UNSUPPORTED;
  if (batch <= 1) { batch = 1; sbytes = ((rows * ld * 4 + 15) / 16) * 16; }
  if (sbytes % 16 != 0 || sbytes <= 0) return CB_ERR_ARG;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 4, (cuuint64_t)sbytes};
  cuuint32_t box[3] = {32u, 32u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CB_OK : CB_ERR_ARG;
}

bool gemm_tc2_supported(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb) {
  return M > 0 && N > 0 && K > 0 && M < (1ll << 30) && N < (1ll << 30) && K < (1ll << 30) && lda % 8 == 0 && ldb % 8 == 0 &&
         lda >= K && ldb >= K && aligned16(A) && aligned16(B);
}

int gemm_tc2(const Gemm2Batch& g, cudaStream_t st) {
  if (!gemm_tc2_supported(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb) || g.batch < 1) return CB_ERR_UNSUPPORTED;
  if (g.batch > 1 && (g.sA % 16 != 0 || g.sB % 16 != 0)) return CB_ERR_UNSUPPORTED;
  Gemm2Args a;
  a.M = (int)g.M; a.N = (int)g.N; a.K = (int)g.K; a.batch = (int)g.batch;
  const int n16 = (int)((g.N + 15) / 16 * 16);
  a.n_tile = n16 < 256 ? n16 : 256;
  if (a.n_tile < 32) a.n_tile = 32;                    // each CTA holds n_tile / 2 >= 16 rows of B (two swizzle atoms)
  a.tiles_m = (int)((g.M + 2 * G2_BM - 1) / (2 * G2_BM));
  a.tiles_n = (int)((g.N + a.n_tile - 1) / a.n_tile);
  a.alpha = g.alpha;
  a.C = g.C; a.ldc = g.ldc; a.sC = g.sC;
  a.Cb = g.Cb; a.ldcb = g.ldcb; a.sCb = g.sCb;
  a.Ct = g.Ct; a.ldct = g.ldct; a.sCt = g.sCt;
  a.colscale = g.colscale; a.sCol = g.sCol; a.rowscale = g.rowscale; a.sRow = g.sRow;
  a.Cs = g.Cs; a.sCs = g.sCs; a.split_mode = g.split_mode;
  a.error_flag = g.error_flag; a.sFlag = 0;
  a.tile_counter = g.tile_counter;
  CUtensorMap ta, tb, tc;
  CB_TRY(make_tmap_bf16_batched(&ta, g.A, g.M, g.K, g.lda, g.batch, g.sA, G2_BM));
  CB_TRY(make_tmap_bf16_batched(&tb, g.B, g.N, g.K, g.ldb, g.batch, g.sB, a.n_tile / 2));
  a.c_tma = 0;
  tc = ta;                                              // (a valid descriptor when the fp32 store path is not taken)
  if (g.C != nullptr && g.Cb == nullptr && g.Ct == nullptr && g.Cs == nullptr && aligned16(g.C) && g.ldc % 4 == 0 &&
      (g.batch <= 1 || g.sC % 16 == 0) && make_tmap_f32_out(&tc, g.C, g.M, g.N, g.ldc, g.batch, g.sC) == CB_OK)
    a.c_tma = 1;
  static PerDeviceOnce once;
  CB_TRY(opt_in_dynamic_smem(gemm_tc2_kernel, G2_SMEM, once));
  const int64_t total_tiles = (int64_t)a.batch * a.tiles_m * a.tiles_n;
  int clusters = g.max_clusters > 0 ? g.max_clusters : kNumSMs / 2;
  if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
  if (total_tiles < clusters) clusters = (int)total_tiles;
  gemm_tc2_kernel<<<2 * clusters, G2_THREADS, G2_SMEM, st>>>(ta, tb, tc, a);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

}  // namespace cb

// C ABI: the batched pair-CTA contraction, exported for validation against a reference GEMM.
extern "C" int cb_gemm_bf16_tn_batched(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const void* A_bf16,
                                       int64_t lda, int64_t stride_a_bytes, const void* B_bf16, int64_t ldb,
                                       int64_t stride_b_bytes, float* C, int64_t ldc, int64_t stride_c_bytes, void* Cb_bf16,
                                       int64_t ldcb, int64_t stride_cb_bytes, void* Ct_bf16, int64_t ldct,
                                       int64_t stride_ct_bytes, const float* colscale, int64_t stride_col_bytes,
                                       const float* rowscale, int64_t stride_row_bytes, int max_clusters, int* tile_counter,
                                       int* error_flag, void* stream) {
  if (A_bf16 == nullptr || B_bf16 == nullptr || (C == nullptr && Cb_bf16 == nullptr && Ct_bf16 == nullptr)) return CB_ERR_ARG;
  cb::Gemm2Batch g;
  g.batch = batch; g.M = M; g.N = N; g.K = K; g.alpha = alpha;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A_bf16); g.lda = lda; g.sA = stride_a_bytes;
  g.B = reinterpret_cast<const __nv_bfloat16*>(B_bf16); g.ldb = ldb; g.sB = stride_b_bytes;
  g.C = C; g.ldc = ldc; g.sC = stride_c_bytes;
  g.Cb = reinterpret_cast<__nv_bfloat16*>(Cb_bf16); g.ldcb = ldcb; g.sCb = stride_cb_bytes;
  g.Ct = reinterpret_cast<__nv_bfloat16*>(Ct_bf16); g.ldct = ldct; g.sCt = stride_ct_bytes;
  g.colscale = colscale; g.sCol = stride_col_bytes; g.rowscale = rowscale; g.sRow = stride_row_bytes;
  g.max_clusters = max_clusters; g.error_flag = error_flag; g.tile_counter = tile_counter;
  return cb::gemm_tc2(g, (cudaStream_t)stream);
}

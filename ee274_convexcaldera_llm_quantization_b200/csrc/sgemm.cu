// fp32 SIMT GEMM with arbitrary element strides on A, B and C.
//
// This is the path for contractions whose shapes the tcgen05 tile kernels (gemm_tc.cu) do
// not cover -- ragged or tiny dimensions, the r x r / q x q products of the least-squares
// and orthonormalisation steps -- and the in-library reference those kernels are checked
// against.  Shared-memory tiled, register micro-tiles; split-K when the output alone cannot
// fill 148 SMs and the caller provides scratch: every K slice parks its register tile in the
// scratch and the last CTA of a tile to arrive sums the slices in slice order (no floating-
// point atomics, so equal inputs give equal bits).
#include "common.cuh"
#include "internal.h"

namespace cb {

template <int MT, bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha,
             const float* __restrict__ A, int64_t a_rs, int64_t a_cs,
             const float* __restrict__ B, int64_t b_rs, int64_t b_cs,
             float* __restrict__ C, int64_t c_rs, int64_t c_cs,
             const float* __restrict__ colscale, int klen, int mode /*0 store, 1 accumulate*/,
             float* __restrict__ split_ws, int64_t bstride /* > 0: blockIdx.z is a batch index, no K split */) {
  if (bstride > 0) {
    const int64_t bo = bstride * blockIdx.z;
    A = boff(A, bo); B = boff(B, bo); C = boff(C, bo); colscale = boff(colscale, bo);
  }
  constexpr int TM = 16 * MT, TN = 16 * MT, TK = 16;
  constexpr int H = MT / 4;
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bm = blockIdx.y * TM, bn = blockIdx.x * TN;
  const int k0 = bstride > 0 ? 0 : blockIdx.z * klen;
  const int k1 = min(K, k0 + klen);
  float acc[MT][MT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) acc[i][j] = 0.f;

  for (int kb = k0; kb < k1; kb += TK) {
#pragma unroll
    for (int i = 0; i < (TM * TK) / 256; ++i) {
      const int e = tid + 256 * i;
      int k, m;
      if (A_KCONTIG) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      const int gm = bm + m, gk = kb + k;
      As[k][m] = (gm < M && gk < k1) ? A[(int64_t)gm * a_rs + (int64_t)gk * a_cs] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (TN * TK) / 256; ++i) {
      const int e = tid + 256 * i;
      int k, n;
      if (B_NCONTIG) { n = e % TN; k = e / TN; } else { k = e % TK; n = e / TK; }
      const int gn = bn + n, gk = kb + k;
      Bs[k][n] = (gn < N && gk < k1) ? B[(int64_t)gk * b_rs + (int64_t)gn * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[MT], b[MT];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 va = *reinterpret_cast<const float4*>(&As[kk][h * 64 + ty * 4]);
        const float4 vb = *reinterpret_cast<const float4*>(&Bs[kk][h * 64 + tx * 4]);
        a[4 * h] = va.x; a[4 * h + 1] = va.y; a[4 * h + 2] = va.z; a[4 * h + 3] = va.w;
        b[4 * h] = vb.x; b[4 * h + 1] = vb.y; b[4 * h + 2] = vb.z; b[4 * h + 3] = vb.w;
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (bstride == 0 && gridDim.z > 1) {
    // split-K: park the raw register tile ([tile][slice][thread][MT*MT]); sgemm_splitk_reduce_kernel
    // follows on the same stream
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    float4* mine = reinterpret_cast<float4*>(split_ws + ((size_t)tile * gridDim.z + blockIdx.z) * (256 * MT * MT) +
                                             (size_t)tid * (MT * MT));
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < MT; j += 4) __stcg(mine + (i * MT + j) / 4, make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]));
    return;
  }

#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int gm = bm + (i / 4) * 64 + ty * 4 + (i & 3);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int gn = bn + (j / 4) * 64 + tx * 4 + (j & 3);
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (colscale != nullptr) v *= colscale[gn];
      float* p = C + (int64_t)gm * c_rs + (int64_t)gn * c_cs;
      if (mode == 1) *p += v;
      else *p = v;
    }
  }
}

// Sums the K slices of a split sgemm in slice order; one thread per 4 consecutive output columns.
template <int MT>
__global__ void __launch_bounds__(256)
sgemm_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int tiles_n, int M, int N, float alpha,
                           float* __restrict__ C, int64_t c_rs, int64_t c_cs, const float* __restrict__ colscale,
                           int mode, int64_t total4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  constexpr int QUADS = MT * MT / 4, T = 16 * MT;
  const int quad = (int)(idx % QUADS);
  const int tid = (int)((idx / QUADS) & 255);
  const int tile = (int)(idx / (QUADS * 256));
  const size_t tile_elems = (size_t)256 * MT * MT;
  const float* src = ws + (size_t)tile * splits * tile_elems + (size_t)tid * (MT * MT) + (size_t)quad * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int z = 0; z < splits; ++z) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * tile_elems));
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  const int i = (quad * 4) / MT, j = (quad * 4) % MT;
  const int tx = tid & 15, ty = tid >> 4;
  const int gm = (tile / tiles_n) * T + (i / 4) * 64 + ty * 4 + (i & 3);
  const int gn0 = (tile % tiles_n) * T + (j / 4) * 64 + tx * 4;
  if (gm >= M) return;
  const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int gn = gn0 + c;
    if (gn >= N) break;
    float x = alpha * v[c];
    if (colscale != nullptr) x *= colscale[gn];
    float* p = C + (int64_t)gm * c_rs + (int64_t)gn * c_cs;
    if (mode == 1) *p += x;
    else *p = x;
  }
}

__global__ void fill_strided_kernel(float* C, int M, int N, int64_t c_rs, int64_t c_cs, float v) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N, c = i - r * N;
    C[r * c_rs + c * c_cs] = v;
  }
}

int sgemm(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t a_rs, int64_t a_cs,
          const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t c_rs, int64_t c_cs,
          bool accumulate, const float* colscale, cudaStream_t st, const SplitWs* sw, const Bt& bt) {
  if (bt.n > 1) sw = nullptr;        // a batch runs unsplit: blockIdx.z carries the batch index
  if (M < 0 || N < 0 || K < 0 || (M > 0 && N > 0 && (C == nullptr)) ||
      (K > 0 && M > 0 && N > 0 && (A == nullptr || B == nullptr)))
    return CB_ERR_ARG;
  if (M == 0 || N == 0) return CB_OK;
  if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return CB_ERR_ARG;
  if (K == 0) {
    if (!accumulate) {
      fill_strided_kernel<<<grid_for(M * N, 256, 4), 256, 0, st>>>(C, (int)M, (int)N, c_rs, c_cs, 0.f);
      CB_CHECK_LAUNCH();
    }
    return CB_OK;
  }
  const bool big = (M >= 512 && N >= 512);
  const int T = big ? 128 : 64;
  const int64_t tiles = ((M + T - 1) / T) * ((N + T - 1) / T);
  int splitk = 1;
  if (tiles < 2 * kNumSMs && K >= 512 && sw != nullptr && sw->buf != nullptr) {
    splitk = (int)((2 * kNumSMs + tiles - 1) / tiles);
    const int maxsplit = (int)(K / 128);
    if (splitk > maxsplit) splitk = maxsplit;
    const int fit = (int)(sw->bytes / ((size_t)tiles * T * T * sizeof(float)));   // scratch the caller provided
    if (splitk > fit) splitk = fit;
    if (splitk < 1) splitk = 1;
  }
  int klen = (int)((K + splitk - 1) / splitk);
  klen = ((klen + 15) / 16) * 16;
  splitk = (int)((K + klen - 1) / klen);
  const int mode = accumulate ? 1 : 0;
  float* split_ws = splitk > 1 ? sw->buf : nullptr;
  dim3 grid((unsigned)((N + T - 1) / T), (unsigned)((M + T - 1) / T), (unsigned)(bt.n > 1 ? bt.n : splitk));
  const int64_t bstride = bt.n > 1 ? bt.stride : 0;
  const bool akc = (a_cs == 1), bnc = (b_cs == 1);
#define CB_SGEMM_LAUNCH(MT, AK, BN)                                                                     \
  sgemm_kernel<MT, AK, BN><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, a_rs, a_cs, B, b_rs, \
                                                 b_cs, C, c_rs, c_cs, colscale, klen, mode, \
                                                 split_ws, bstride)
  if (big) {
    if (akc && bnc) CB_SGEMM_LAUNCH(8, true, true);
    else if (akc) CB_SGEMM_LAUNCH(8, true, false);
    else if (bnc) CB_SGEMM_LAUNCH(8, false, true);
    else CB_SGEMM_LAUNCH(8, false, false);
  } else {
    if (akc && bnc) CB_SGEMM_LAUNCH(4, true, true);
    else if (akc) CB_SGEMM_LAUNCH(4, true, false);
    else if (bnc) CB_SGEMM_LAUNCH(4, false, true);
    else CB_SGEMM_LAUNCH(4, false, false);
  }
#undef CB_SGEMM_LAUNCH
  CB_CHECK_LAUNCH();
  if (splitk > 1) {
    const int64_t total4 = tiles * 256 * (big ? 16 : 4);
    const unsigned blocks = (unsigned)((total4 + 255) / 256);
    const int tiles_n = (int)((N + T - 1) / T);
    if (big)
      sgemm_splitk_reduce_kernel<8><<<blocks, 256, 0, st>>>(split_ws, splitk, tiles_n, (int)M, (int)N, alpha, C, c_rs, c_cs,
                                                           colscale, mode, total4);
    else
      sgemm_splitk_reduce_kernel<4><<<blocks, 256, 0, st>>>(split_ws, splitk, tiles_n, (int)M, (int)N, alpha, C, c_rs, c_cs,
                                                           colscale, mode, total4);
    CB_CHECK_LAUNCH();
  }
  return CB_OK;
}

}  // namespace cb

extern "C" int cb_sgemm_strided(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t a_rs,
                                int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                                int64_t c_rs, int64_t c_cs, int accumulate, void* stream) {
  return cb::sgemm(M, N, K, alpha, A, a_rs, a_cs, B, b_rs, b_cs, C, c_rs, c_cs, accumulate != 0, nullptr,
                   (cudaStream_t)stream, nullptr);
}

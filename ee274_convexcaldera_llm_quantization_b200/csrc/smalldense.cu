// Small dense factorizations that sit between the big contractions:
//   * blocked Cholesky + triangular inverse of a q x q Gram matrix (q <= 512), used by the
//     CholeskyQR re-orthonormalisation of the sketch and by the normal-equation solves of
//     the LPLR updates (replaces torch.linalg.lstsq's QR, alg.py:163,175);
//   * one-sided Jacobi eigensolver on the Cholesky factor, the Rayleigh-Ritz step that
//     replaces the truncated SVD of LR_init (alg.py:214-217).
// Both are single-CTA, latency-bound kernels: the matrices are at most 1 MiB and stay in
// L1/L2; panels are staged in shared memory.
#include <cooperative_groups.h>
#include "common.cuh"
#include "internal.h"

namespace cg = cooperative_groups;

namespace cb {

constexpr int NB = 32;          // panel width
constexpr int QMAX = 512;       // largest supported matrix
constexpr int TMAXR = QMAX - NB;  // most rows below a diagonal block

struct CholSmem {
  float D[NB][NB + 4];             // diagonal block / its Cholesky factor (rows 16-byte aligned)
  float Praw[TMAXR][NB + 1];       // scratch shared with the triangular-inverse chains
  float Pt[NB][TMAXR + 4];         // solved panel, transposed: Pt[c][row]
  float diag0[QMAX];               // backup of the diagonal for the ridge retry
  float rdiag[NB];                 // reciprocal diagonal of the current 32 x 32 factor
  int fail;
  float mean_diag;
};

// In-place Cholesky of the lower triangle of G (row-major, leading dimension q).  The strict
// upper triangle is never written, so together with diag0 it is a backup of the input.
// Per 32-wide panel: (1) warp 0 factors the diagonal block in registers (lane = row, all
// indices static after unrolling, broadcasts by shuffle); (2) every thread owns one row below
// the block and solves x L_kk^T = a by forward substitution in registers, reading L_kk as
// 128-bit shared-memory broadcasts; (3) 4 x 4 register tiles apply the rank-32 update to the
// trailing lower triangle from the transposed panel in shared memory.
__device__ void chol_factor(float* __restrict__ G, int q, CholSmem& s, long long* stamps = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < q; k0 += NB) {
    const int nb = min(NB, q - k0);
    long long* ps = (stamps != nullptr && tid == 0 && k0 / NB < 3) ? stamps + 4 + 5 * (k0 / NB) : nullptr;
    if (ps) ps[0] = clock64();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int i = e >> 5, j = e & 31;
      float v = 0.f;
      if (i < nb && j <= i) v = G[(size_t)(k0 + i) * q + k0 + j];
      s.D[i][j] = v;
    }
    __syncthreads();
    if (ps) ps[1] = clock64();   // diagonal block in shared memory
    if (warp == 0) {
      float a[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = s.D[lane][c];
      bool bad = false;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float d = __shfl_sync(0xffffffffu, a[j], j);
        if (j >= nb) d = 1.f;
        if (!(d > 0.f) || d > 3.0e38f) { bad = true; d = 1.f; }   // non-positive, NaN or inf pivot
        const float inv = rsqrtf(d), sd = d * inv;                   // 2-ulp rsqrt is ample for a Gram factor
        if (lane == 0) s.rdiag[j] = inv;                              // 1 / L[j][j]
        const float lij = (lane > j) ? a[j] * inv : (lane == j ? sd : 0.f);
        a[j] = (lane >= j) ? lij : 0.f;
        // constant trip counts on both loops so the unroller resolves every index statically
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          if (c > j) {
            const float lcj = __shfl_sync(0xffffffffu, lij, c);   // L[c][j]
            if (lane >= c) a[c] = fmaf(-lij, lcj, a[c]);
          }
        }
      }
      if (bad && lane == 0) s.fail = 1;
#pragma unroll
      for (int c = 0; c < NB; ++c) s.D[lane][c] = a[c];
    }
    __syncthreads();
    if (ps) ps[2] = clock64();   // diagonal block factored
    // factor back to global
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int i = e >> 5, j = e & 31;
      if (i < nb && j <= i) G[(size_t)(k0 + i) * q + k0 + j] = s.D[i][j];
    }
    const int T = q - k0 - nb;  // rows below the block
    if (T <= 0) { __syncthreads(); continue; }
    // panel: one row per thread (T <= 480 < blockDim), x <- x L_kk^-T by forward substitution
    for (int row = tid; row < T; row += blockDim.x) {
      float* grow = G + (size_t)(k0 + nb + row) * q + k0;
      float x[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? grow[c] : 0.f;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        float acc = x[c];
#pragma unroll
        for (int k4 = 0; k4 < NB / 4; ++k4) {
          if (4 * k4 < c) {
            const float4 d = *reinterpret_cast<const float4*>(&s.D[c][4 * k4]);   // broadcast
            if (4 * k4 + 0 < c) acc = fmaf(-x[4 * k4 + 0], d.x, acc);
            if (4 * k4 + 1 < c) acc = fmaf(-x[4 * k4 + 1], d.y, acc);
            if (4 * k4 + 2 < c) acc = fmaf(-x[4 * k4 + 2], d.z, acc);
            if (4 * k4 + 3 < c) acc = fmaf(-x[4 * k4 + 3], d.w, acc);
          }
        }
        x[c] = acc * s.rdiag[c];
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < nb) grow[c] = x[c];
        s.Pt[c][row] = (c < nb) ? x[c] : 0.f;
      }
    }
    // zero-pad the panel to a multiple of 4 rows so 128-bit reads below stay in bounds
    for (int e = tid; e < 4 * NB; e += blockDim.x) {
      const int c = e >> 2, i = T + (e & 3);
      if (i < TMAXR + 4) s.Pt[c][i] = 0.f;
    }
    __syncthreads();
    if (ps) ps[3] = clock64();   // panel solved
    // trailing update on the lower triangle, 4x4 register tiles
    const int Ti = (T + 3) >> 2;
    const int ntile = Ti * (Ti + 1) / 2;
    for (int t = tid; t < ntile; t += blockDim.x) {
      int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      const int tj = t - ti * (ti + 1) / 2;
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
      for (int c = 0; c < NB; ++c) {
        const float4 pa = *reinterpret_cast<const float4*>(&s.Pt[c][ti * 4]);
        const float4 pb = *reinterpret_cast<const float4*>(&s.Pt[c][tj * 4]);
        const float av[4] = {pa.x, pa.y, pa.z, pa.w}, bv[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
      // read-modify-write of the 4 x 4 tile: all loads first (the compiler must keep a load behind an earlier
      // store to G, which would serialise sixteen L2 round trips)
      float g[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ti * 4 + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = tj * 4 + b;
          g[a][b] = (i < T && j <= i) ? G[(size_t)(k0 + nb + i) * q + k0 + nb + j] : 0.f;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ti * 4 + a;
        if (i >= T) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = tj * 4 + b;
          if (j > i) continue;
          G[(size_t)(k0 + nb + i) * q + k0 + nb + j] = g[a][b] - acc[a][b];
        }
      }
    }
    __syncthreads();
    if (ps) ps[4] = clock64();   // trailing update done
  }
}

// Inverses of all 32 x 32 diagonal blocks of the factor at once: warp b takes block b, lane =
// row while loading / column of the inverse while solving, L[i][k] is broadcast from lane i.
// Every routine that writes an element of Linv also writes its bf16 copy when `Lb` is given (the
// tensor-core contractions consume the inverse as a bf16 operand), so no separate conversion pass runs.
__device__ void diag_block_inverses(const float* G, int q, float* Linv, __nv_bfloat16* Lb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nblk = (q + NB - 1) / NB;
  for (int b = warp; b < nblk; b += nwarps) {
    const int r0 = b * NB, nb = min(NB, q - r0);
    float a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c)
      a[c] = (lane < nb && c <= lane) ? G[(size_t)(r0 + lane) * q + r0 + c] : (c == lane ? 1.f : 0.f);
    float x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      float acc = (lane == i) ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        if (k < i) {
          const float lik = __shfl_sync(0xffffffffu, a[k], i);
          acc = fmaf(-lik, x[k], acc);
        }
      }
      const float lii = __shfl_sync(0xffffffffu, a[i], i);
      x[i] = (lane <= i) ? acc / lii : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i < nb && lane <= i && lane < nb) {
        const size_t e = (size_t)(r0 + i) * q + r0 + lane;
        Linv[e] = x[i];
        if (Lb != nullptr) Lb[e] = __float2bfloat16_rn(x[i]);
      }
  }
}

// Linv = Lc^-1 by block forward substitution; diagonal blocks already hold their inverses.
// Block columns are independent, so up to 8 "chains" of 128 threads each walk one block
// column at a time (chain c takes columns c and nblk-1-c, which balances the triangular
// work); every 32x32x32 block product is staged through chain-private shared tiles.
//   X[i][j] = -Dinv_i * sum_{k=j}^{i-1} L[i][k] X[k][j]
struct ChainTiles {
  float Lt[NB][NB + 1];   // L[i][k] tile, later the partial sum S
  float Xt[NB][NB + 4];   // X[k][j] tile, later Dinv_i (16-byte aligned rows)
};

__device__ __forceinline__ void chain_sync(int chain) {
  asm volatile("bar.sync %0, 128;" ::"r"(chain + 1) : "memory");
}

__device__ void tri_inverse(const float* G, int q, float* Linv, __nv_bfloat16* Lb, ChainTiles* tiles) {
  const int tid = threadIdx.x;
  const int nblk = (q + NB - 1) / NB;
  // strict upper triangle and not-yet-computed blocks start at zero
  for (int e = tid; e < q * q; e += blockDim.x) {
    const int i = e / q, j = e - i * q;
    if ((i / NB) != (j / NB) || j > i) {
      Linv[e] = 0.f;
      if (Lb != nullptr) Lb[e] = __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  const int chain = tid >> 7, ct = tid & 127;
  const int nchains = min((int)(blockDim.x >> 7), (nblk + 1) / 2);
  if (chain < nchains) {
    ChainTiles& T = tiles[chain];
    const int orow = ct >> 2, oc0 = (ct & 3) * 8;     // this thread's 1 x 8 strip of a 32 x 32 block
    for (int pass = 0; pass < 2; ++pass) {
      for (int jj = chain; jj < (nblk + 1) / 2; jj += nchains) {
        const int j = pass == 0 ? jj : nblk - 1 - jj;
        if (pass == 1 && j == jj) continue;            // middle column handled in pass 0
        for (int i = j + 1; i < nblk; ++i) {
          float acc[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = 0.f;
          for (int k = j; k < i; ++k) {
            chain_sync(chain);
            for (int e = ct; e < NB * NB; e += 128) {
              const int r = e >> 5, c = e & 31;
              const int gr = i * NB + r, gc = k * NB + c;
              T.Lt[r][c] = (gr < q && gc < q) ? G[(size_t)gr * q + gc] : 0.f;
              const int xr = k * NB + r, xc = j * NB + c;
              T.Xt[r][c] = (xr < q && xc < q) ? Linv[(size_t)xr * q + xc] : 0.f;
            }
            chain_sync(chain);
#pragma unroll 8
            for (int kk = 0; kk < NB; ++kk) {
              const float l = T.Lt[orow][kk];
              const float4 x0 = *reinterpret_cast<const float4*>(&T.Xt[kk][oc0]);
              const float4 x1 = *reinterpret_cast<const float4*>(&T.Xt[kk][oc0 + 4]);
              acc[0] = fmaf(l, x0.x, acc[0]); acc[1] = fmaf(l, x0.y, acc[1]);
              acc[2] = fmaf(l, x0.z, acc[2]); acc[3] = fmaf(l, x0.w, acc[3]);
              acc[4] = fmaf(l, x1.x, acc[4]); acc[5] = fmaf(l, x1.y, acc[5]);
              acc[6] = fmaf(l, x1.z, acc[6]); acc[7] = fmaf(l, x1.w, acc[7]);
            }
          }
          // X[i][j] = -Dinv_i * S
          chain_sync(chain);
#pragma unroll
          for (int c = 0; c < 8; ++c) T.Lt[orow][oc0 + c] = acc[c];
          for (int e = ct; e < NB * NB; e += 128) {
            const int r = e >> 5, c = e & 31;
            const int gr = i * NB + r, gc = i * NB + c;
            T.Xt[r][c] = (gr < q && gc < q && c <= r) ? Linv[(size_t)gr * q + gc] : 0.f;
          }
          chain_sync(chain);
          float out[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) out[c] = 0.f;
          for (int kk = 0; kk <= orow; ++kk) {
            const float d = T.Xt[orow][kk];
#pragma unroll
            for (int c = 0; c < 8; ++c) out[c] = fmaf(d, T.Lt[kk][oc0 + c], out[c]);
          }
          const int gr = i * NB + orow;
          if (gr < q) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int gc = j * NB + oc0 + c;
              if (gc < q) {
                Linv[(size_t)gr * q + gc] = -out[c];
                if (Lb != nullptr) Lb[(size_t)gr * q + gc] = __float2bfloat16_rn(-out[c]);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
}

constexpr int CHOL_THREADS = 512;   // 128 registers/thread: the 32x32 diagonal factor lives in registers

#ifdef CB_MEASURE
long long* g_chol_timing = nullptr;   // measurement aid (cb_set_chol_timing): 4 clock64 stamps
#else
constexpr long long* g_chol_timing = nullptr;
#endif

__global__ void __launch_bounds__(CHOL_THREADS, 1)
chol_inv_kernel(float* __restrict__ G_, int q, float* __restrict__ Linv_, __nv_bfloat16* __restrict__ Linv_bf16_,
                int* __restrict__ status_, long long* __restrict__ stamps, int64_t bstride) {
  // one CTA per layer of a batch (blockIdx.x); layer b's buffers sit b * bstride bytes after layer 0's
  const int64_t bo = bstride * blockIdx.x;
  float* __restrict__ G = boff(G_, bo);
  float* __restrict__ Linv = boff(Linv_, bo);
  __nv_bfloat16* __restrict__ Linv_bf16 = boff(Linv_bf16_, bo);
  int* __restrict__ status = boff(status_, bo);
  if (blockIdx.x != 0) stamps = nullptr;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem& s = *reinterpret_cast<CholSmem*>(smem_raw);
  const int tid = threadIdx.x;
  double dsum = 0.0;
  for (int i = tid; i < q; i += blockDim.x) { s.diag0[i] = G[(size_t)i * q + i]; }
  if (tid == 0) s.fail = 0;
  __syncthreads();
  if (tid < 32) {
    for (int i = tid; i < q; i += 32) dsum += (double)s.diag0[i];
    dsum = warp_sum(dsum);
    if (tid == 0) s.mean_diag = (float)(dsum / (double)q);
  }
  __syncthreads();
  int retries = 0;
  if (stamps != nullptr && tid == 0) stamps[0] = clock64();

  while (true) {
    chol_factor(G, q, s, stamps);
    __syncthreads();
    const int failed = s.fail;
    __syncthreads();
    if (!failed || retries >= 3) { if (failed) retries = 99; break; }
    // restore the lower triangle from the untouched upper one, add a ridge, retry
    const float ridge = (retries == 0 ? 1e-6f : (retries == 1 ? 1e-4f : 1e-2f)) * fmaxf(s.mean_diag, 1e-30f);
    for (int e = tid; e < q * q; e += blockDim.x) {
      const int i = e / q, j = e - i * q;
      if (j < i) G[e] = G[(size_t)j * q + i];
      else if (j == i) G[e] = s.diag0[i] + ridge;
    }
    if (tid == 0) s.fail = 0;
    ++retries;
    __syncthreads();
  }
  if (stamps != nullptr && tid == 0) stamps[1] = clock64();      // factor done
  if (Linv != nullptr) {
    diag_block_inverses(G, q, Linv, Linv_bf16);
    __syncthreads();
    if (stamps != nullptr && tid == 0) stamps[2] = clock64();    // diagonal block inverses done
    tri_inverse(G, q, Linv, Linv_bf16, reinterpret_cast<ChainTiles*>(&s.Praw[0][0]));
    if (stamps != nullptr && tid == 0) stamps[3] = clock64();    // inverse done
  }
  if (tid == 0 && status != nullptr) atomicMax(status, retries);
}

int cholesky_inverse(float* G, int q, float* Linv, int* status, cudaStream_t st, __nv_bfloat16* Linv_bf16, const Bt& bt) {
  if (G == nullptr || q <= 0) return CB_ERR_ARG;
  if (q > QMAX) return CB_ERR_UNSUPPORTED;
  if (debug_skip() & 1) return CB_OK;
  static PerDeviceOnce once;
  CB_TRY(opt_in_dynamic_smem(chol_inv_kernel, (int)sizeof(CholSmem), once));
  chol_inv_kernel<<<bt.n, CHOL_THREADS, sizeof(CholSmem), st>>>(G, q, Linv, Linv_bf16, status, g_chol_timing, bt.stride);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// ---------------------------------------------------------------- one-sided Jacobi
// Orthogonalises the columns of Lc (G = Lc Lc^T) by plane rotations; on exit the columns
// are U * Sigma, so eigenvalues of G are squared column norms and eigenvectors the
// normalised columns.  Columns are kept as contiguous rows of `work`.
constexpr int JMAXV = QMAX / 32;  // elements of one vector held per lane

__global__ void __launch_bounds__(1024, 1)
jacobi_kernel(const float* __restrict__ Lc, int q, float* __restrict__ evals, float* __restrict__ evecs,
              float* __restrict__ work, int* __restrict__ sweeps_out, int max_sweeps, float tol) {
  __shared__ int s_rot;
  __shared__ float s_lam[QMAX];
  __shared__ int s_rank[QMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // work[v][i] = Lc[i][v] (i >= v), 0 above the diagonal
  for (int e = tid; e < q * q; e += blockDim.x) {
    const int i = e / q, v = e - i * q;
    work[(size_t)v * q + i] = (i >= v) ? Lc[e] : 0.f;
  }
  __syncthreads();
  const int qe = q + (q & 1);
  const int npairs = qe >> 1;
  const int nper = (q + 31) >> 5;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int t = 0; t < qe - 1; ++t) {
      for (int pi = warp; pi < npairs; pi += nwarps) {
        int a, b;
        if (pi == 0) { a = qe - 1; b = t; }
        else { a = (t + pi) % (qe - 1); b = (t - pi + (qe - 1)) % (qe - 1); }
        if (a >= q || b >= q) continue;
        float* xa = work + (size_t)a * q;
        float* xb = work + (size_t)b * q;
        float x[JMAXV], y[JMAXV];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int k = 0; k < JMAXV; ++k) {
          const int i = lane + 32 * k;
          x[k] = 0.f; y[k] = 0.f;
          if (k < nper && i < q) { x[k] = xa[i]; y[k] = xb[i]; }
          al = fmaf(x[k], x[k], al);
          be = fmaf(y[k], y[k], be);
          ga = fmaf(x[k], y[k], ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        if (fabsf(ga) > tol * sqrtf(al * be) && al > 0.f && be > 0.f) {
          // rotation parameters in fp64: c and s are then rounded independently, so c^2 + s^2 - 1
          // has no systematic sign and column norms do not drift over the ~q rotations per sweep
          const double zeta = ((double)be - (double)al) / (2.0 * (double)ga);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cd = 1.0 / sqrt(1.0 + tt * tt);
          const float c = (float)cd, sn = (float)(cd * tt);
#pragma unroll
          for (int k = 0; k < JMAXV; ++k) {
            const int i = lane + 32 * k;
            if (k < nper && i < q) {
              xa[i] = c * x[k] - sn * y[k];
              xb[i] = sn * x[k] + c * y[k];
            }
          }
          if (lane == 0) atomicAdd(&s_rot, 1);
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (rot * 16 < q) { ++sweep; break; }
  }
  // eigenvalues = squared norms
  for (int v = warp; v < q; v += nwarps) {
    float acc = 0.f;
    for (int i = lane; i < q; i += 32) { const float z = work[(size_t)v * q + i]; acc = fmaf(z, z, acc); }
    acc = warp_sum(acc);
    if (lane == 0) s_lam[v] = acc;
  }
  __syncthreads();
  for (int v = tid; v < q; v += blockDim.x) {
    const float lv = s_lam[v];
    int rk = 0;
    for (int u = 0; u < q; ++u) {
      const float lu = s_lam[u];
      rk += (lu > lv) || (lu == lv && u < v);
    }
    s_rank[v] = rk;
  }
  __syncthreads();
  for (int v = warp; v < q; v += nwarps) {
    const float lam = s_lam[v];
    const float inv = lam > 0.f ? __fdiv_rn(1.f, __fsqrt_rn(lam)) : 0.f;
    const int rk = s_rank[v];
    for (int i = lane; i < q; i += 32) evecs[(size_t)rk * q + i] = work[(size_t)v * q + i] * inv;
    if (lane == 0) evals[rk] = lam;
  }
  if (tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep;
}

// Shared-memory variant for q <= 224 (q % 4 == 0): the q x q working matrix (<= 204 KB) lives
// in shared memory for the whole solve, so a round costs one pass of shared-memory traffic
// instead of a round trip to L2.  512 threads = 64 groups of 8 lanes; a group owns one
// column pair at a time and keeps both columns in registers between the dot products and
// the rotation (one load + one store per element and round).
constexpr int JS_QMAX = 224;
constexpr int JS_THREADS = 512;
constexpr int JS_V4 = JS_QMAX / 32;   // float4 per lane and vector

__global__ void __launch_bounds__(JS_THREADS, 1)
jacobi_smem_kernel(const float* __restrict__ Lc_, int q, float* __restrict__ evals_, float* __restrict__ evecs_,
                   int* __restrict__ sweeps_out_, int max_sweeps, float tol, int64_t bstride) {
  // one CTA per layer of a batch (blockIdx.x)
  const int64_t bo = bstride * blockIdx.x;
  const float* __restrict__ Lc = boff(Lc_, bo);
  float* __restrict__ evals = boff(evals_, bo);
  float* __restrict__ evecs = boff(evecs_, bo);
  int* __restrict__ sweeps_out = boff(sweeps_out_, bo);
  extern __shared__ __align__(16) float js_smem[];
  const int stride = q + 4;
  float* V = js_smem;                       // V[v * stride + i] = component i of vector v
  float* s_lam = V + (size_t)q * stride;    // q
  int* s_rank = reinterpret_cast<int*>(s_lam + q);
  __shared__ int s_rot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = JS_THREADS >> 5;
  const int g = tid >> 3, gl = tid & 7, ngroups = JS_THREADS >> 3;
  for (int e = tid; e < q * q; e += JS_THREADS) {
    const int i = e / q, v = e - i * q;
    V[v * stride + i] = (i >= v) ? Lc[e] : 0.f;
  }
  __syncthreads();
  const int qe = q + (q & 1);
  const int npairs = qe >> 1;
  const int nv4 = (q + 31) >> 5;            // float4 per lane (8 lanes x 4 floats = 32 elements per step)
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int t = 0; t < qe - 1; ++t) {
      // trip count is uniform across the CTA: the group shuffles below use the full warp mask
      for (int pbase = 0; pbase < npairs; pbase += ngroups) {
        const int pi = pbase + g;
        int a = 0, b = 0;
        if (pi == 0) { a = qe - 1; b = t; }
        else if (pi < npairs) { a = (t + pi) % (qe - 1); b = (t - pi + (qe - 1)) % (qe - 1); }
        const bool live = (pi < npairs && a < q && b < q);
        float4* xa = reinterpret_cast<float4*>(V + (size_t)(live ? a : 0) * stride);
        float4* xb = reinterpret_cast<float4*>(V + (size_t)(live ? b : 0) * stride);
        float4 x[JS_V4], y[JS_V4];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int k = 0; k < JS_V4; ++k) {
          const int i4 = gl + 8 * k;         // float4 index; element 4*i4
          if (k < nv4 && 4 * i4 < q && live) { x[k] = xa[i4]; y[k] = xb[i4]; }
          else { x[k] = make_float4(0.f, 0.f, 0.f, 0.f); y[k] = x[k]; }
          al = fmaf(x[k].x, x[k].x, al); al = fmaf(x[k].y, x[k].y, al); al = fmaf(x[k].z, x[k].z, al); al = fmaf(x[k].w, x[k].w, al);
          be = fmaf(y[k].x, y[k].x, be); be = fmaf(y[k].y, y[k].y, be); be = fmaf(y[k].z, y[k].z, be); be = fmaf(y[k].w, y[k].w, be);
          ga = fmaf(x[k].x, y[k].x, ga); ga = fmaf(x[k].y, y[k].y, ga); ga = fmaf(x[k].z, y[k].z, ga); ga = fmaf(x[k].w, y[k].w, ga);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        const bool rotate = live && fabsf(ga) > tol * sqrtf(al * be) && al > 0.f && be > 0.f;
        // rotation parameters in fp64: c and s are then rounded independently, so c^2 + s^2 - 1
        // has no systematic sign and column norms do not drift over the ~q rotations per sweep.
        // One lane per group runs the fp64 divide / square-root sequences and the others get the result
        // by shuffle (same bits as computing it on every lane; measured neutral in time: the round is
        // bound by streaming the matrix through shared memory).
        float c = 1.f, sn = 0.f;
        if (rotate && gl == 0) {
          const double zeta = ((double)be - (double)al) / (2.0 * (double)ga);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cd = 1.0 / sqrt(1.0 + tt * tt);
          c = (float)cd; sn = (float)(cd * tt);
        }
        c = __shfl_sync(0xffffffffu, c, lane & ~7);
        sn = __shfl_sync(0xffffffffu, sn, lane & ~7);
        if (rotate) {
#pragma unroll
          for (int k = 0; k < JS_V4; ++k) {
            const int i4 = gl + 8 * k;
            if (k < nv4 && 4 * i4 < q) {
              xa[i4] = make_float4(c * x[k].x - sn * y[k].x, c * x[k].y - sn * y[k].y, c * x[k].z - sn * y[k].z, c * x[k].w - sn * y[k].w);
              xb[i4] = make_float4(sn * x[k].x + c * y[k].x, sn * x[k].y + c * y[k].y, sn * x[k].z + c * y[k].z, sn * x[k].w + c * y[k].w);
            }
          }
          if (gl == 0) atomicAdd(&s_rot, 1);
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (rot * 16 < q) { ++sweep; break; }
  }
  for (int v = warp; v < q; v += nwarps) {
    float acc = 0.f;
    for (int i = lane; i < q; i += 32) { const float z = V[v * stride + i]; acc = fmaf(z, z, acc); }
    acc = warp_sum(acc);
    if (lane == 0) s_lam[v] = acc;
  }
  __syncthreads();
  for (int v = tid; v < q; v += JS_THREADS) {
    const float lv = s_lam[v];
    int rk = 0;
    for (int u = 0; u < q; ++u) {
      const float lu = s_lam[u];
      rk += (lu > lv) || (lu == lv && u < v);
    }
    s_rank[v] = rk;
  }
  __syncthreads();
  for (int v = warp; v < q; v += nwarps) {
    const float lam = s_lam[v];
    const float inv = lam > 0.f ? __fdiv_rn(1.f, __fsqrt_rn(lam)) : 0.f;
    const int rk = s_rank[v];
    for (int i = lane; i < q; i += 32) evecs[(size_t)rk * q + i] = V[v * stride + i] * inv;
    if (lane == 0) evals[rk] = lam;
  }
  if (tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep;
}

// Thread-block-cluster variant: 8 CTAs (one cluster, 8 SMs) share every round of the
// round-robin tournament.  The q x q working matrix lives in global memory and stays in L2
// (all accesses bypass L1: ld/st.global.cg), a group of LPP lanes owns one column pair per
// round and keeps both columns in registers between the dot products and the rotation, and
// rounds are separated by a hardware cluster barrier (release/acquire at cluster scope).
// LPP = 8 lanes per pair for q <= 256 (256 threads per CTA), 16 for q <= 512 (512 threads).
constexpr int JC_CTAS = 8;

template <int LPP, int NV4>
__global__ void __cluster_dims__(JC_CTAS, 1, 1) __launch_bounds__(32 * LPP, 1)
jacobi_cluster_kernel(const float* __restrict__ Lc, int q, float* __restrict__ evals, float* __restrict__ evecs,
                      float* work, int* sweeps_out, int* rot_counters, float* lam_buf, int max_sweeps, float tol) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int s_rot;
  const int cta = (int)cluster.block_rank();
  const int nthreads = 32 * LPP;
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / LPP, gl = tid % LPP;
  const int groups_per_cta = nthreads / LPP;           // 32
  const int NG = groups_per_cta * JC_CTAS;             // 256 pairs per pass
  const int ggroup = cta * groups_per_cta + g;
  // work[v][i] = Lc[i][v] (i >= v), 0 above the diagonal
  for (int e = cta * nthreads + tid; e < q * q; e += nthreads * JC_CTAS) {
    const int i = e / q, v = e - i * q;
    __stcg(work + (size_t)v * q + i, (i >= v) ? Lc[e] : 0.f);
  }
  if (tid == 0) s_rot = 0;
  cluster.sync();
  const int qe = q + (q & 1);
  const int npairs = qe >> 1;
  const int q4 = q >> 2;                               // q % 4 == 0
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    for (int t = 0; t < qe - 1; ++t) {
      for (int pbase = 0; pbase < npairs; pbase += NG) {
        const int pi = pbase + ggroup;
        int a = 0, b = 0;
        if (pi == 0) { a = qe - 1; b = t; }
        else if (pi < npairs) { a = (t + pi) % (qe - 1); b = (t - pi + (qe - 1)) % (qe - 1); }
        const bool live = (pi < npairs && a < q && b < q);
        float4* xa = reinterpret_cast<float4*>(work + (size_t)(live ? a : 0) * q);
        float4* xb = reinterpret_cast<float4*>(work + (size_t)(live ? b : 0) * q);
        float4 x[NV4], y[NV4];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
          const int i4 = gl + LPP * k;
          if (i4 < q4 && live) { x[k] = __ldcg(xa + i4); y[k] = __ldcg(xb + i4); }
          else { x[k] = make_float4(0.f, 0.f, 0.f, 0.f); y[k] = x[k]; }
        }
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
          al = fmaf(x[k].x, x[k].x, al); al = fmaf(x[k].y, x[k].y, al); al = fmaf(x[k].z, x[k].z, al); al = fmaf(x[k].w, x[k].w, al);
          be = fmaf(y[k].x, y[k].x, be); be = fmaf(y[k].y, y[k].y, be); be = fmaf(y[k].z, y[k].z, be); be = fmaf(y[k].w, y[k].w, be);
          ga = fmaf(x[k].x, y[k].x, ga); ga = fmaf(x[k].y, y[k].y, ga); ga = fmaf(x[k].z, y[k].z, ga); ga = fmaf(x[k].w, y[k].w, ga);
        }
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        const bool rotate = live && fabsf(ga) > tol * sqrtf(al * be) && al > 0.f && be > 0.f;
        float c = 1.f, sn = 0.f;
        if (rotate && gl == 0) {        // one lane per group runs the fp64 sequences, see jacobi_smem_kernel
          const double zeta = ((double)be - (double)al) / (2.0 * (double)ga);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cd = 1.0 / sqrt(1.0 + tt * tt);
          c = (float)cd; sn = (float)(cd * tt);
        }
        c = __shfl_sync(0xffffffffu, c, lane & ~(LPP - 1));
        sn = __shfl_sync(0xffffffffu, sn, lane & ~(LPP - 1));
        if (rotate) {
#pragma unroll
          for (int k = 0; k < NV4; ++k) {
            const int i4 = gl + LPP * k;
            if (i4 < q4) {
              __stcg(xa + i4, make_float4(c * x[k].x - sn * y[k].x, c * x[k].y - sn * y[k].y, c * x[k].z - sn * y[k].z, c * x[k].w - sn * y[k].w));
              __stcg(xb + i4, make_float4(sn * x[k].x + c * y[k].x, sn * x[k].y + c * y[k].y, sn * x[k].z + c * y[k].z, sn * x[k].w + c * y[k].w));
            }
          }
          if (gl == 0) atomicAdd(&s_rot, 1);
        }
      }
      cluster.sync();
      // the other parity's counter was last read right after the previous sweep's final barrier;
      // every CTA is past that read once it has arrived at this round's barrier
      if (t == 0 && cta == 0 && tid == 0) __stcg(rot_counters + ((sweep + 1) & 1), 0);
    }
    if (tid == 0) {
      if (s_rot > 0) atomicAdd(rot_counters + (sweep & 1), s_rot);
      s_rot = 0;
    }
    cluster.sync();
    const int rot = __ldcg(rot_counters + (sweep & 1));
    // quadratic convergence: once fewer than q/16 pairs still exceed the (3e-5) threshold, what is left
    // is at rounding level of the captured energy; do not spend a full confirmation sweep on it
    if (rot * 16 < q) { ++sweep; break; }
  }
  // eigenvalues = squared column norms
  for (int v = ggroup; v < q; v += NG) {
    const float4* xv = reinterpret_cast<const float4*>(work + (size_t)v * q);
    float acc = 0.f;
    for (int i4 = gl; i4 < q4; i4 += LPP) { const float4 z = __ldcg(xv + i4); acc = fmaf(z.x, z.x, acc); acc = fmaf(z.y, z.y, acc); acc = fmaf(z.z, z.z, acc); acc = fmaf(z.w, z.w, acc); }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (gl == 0) __stcg(lam_buf + v, acc);
  }
  cluster.sync();
  // rank (descending, ties by index) and normalised, sorted output; one group per vector
  for (int vbase = 0; vbase < q; vbase += NG) {
    const int v = vbase + ggroup;
    const bool live = v < q;
    const float lv = live ? __ldcg(lam_buf + v) : 0.f;
    int rk = 0;
    if (live)
      for (int u = gl; u < q; u += LPP) { const float lu = __ldcg(lam_buf + u); rk += (lu > lv) || (lu == lv && u < v); }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    if (live) {
      const float inv = lv > 0.f ? __fdiv_rn(1.f, __fsqrt_rn(lv)) : 0.f;
      const float4* xv = reinterpret_cast<const float4*>(work + (size_t)v * q);
      float4* ov = reinterpret_cast<float4*>(evecs + (size_t)rk * q);
      for (int i4 = gl; i4 < q4; i4 += LPP) { float4 z = __ldcg(xv + i4); z.x *= inv; z.y *= inv; z.z *= inv; z.w *= inv; ov[i4] = z; }
      if (gl == 0) evals[rk] = lv;
    }
  }
  if (cta == 0 && tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep;
}

// A column pair is rotated while |<x,y>| > tol * |x| |y|.  The captured energy of the
// Rayleigh-Ritz step is second order in the residual coupling, so 3e-5 leaves it exact to fp32.
constexpr float kJacobiTol = 3e-5f;


int jacobi_eigh_from_chol(const float* Lc, int q, float* evals, float* evecs, float* work, int* sweeps,
                          cudaStream_t st, const Bt& bt) {
  if (Lc == nullptr || evals == nullptr || evecs == nullptr || work == nullptr || q <= 0) return CB_ERR_ARG;
  if (q > QMAX) return CB_ERR_UNSUPPORTED;
  if (debug_skip() & 2) return CB_OK;
  // The 8-CTA cluster kernel has the lower latency (one layer in flight); the single-CTA kernel spends
  // ~3x less SM time (many layers in flight).  The caller's execution policy selects (cb_caldera_params.exec_mode).
  if (bt.n > 1 && !(q <= JS_QMAX && q % 4 == 0)) return CB_ERR_UNSUPPORTED;   // batches run the one-CTA-per-layer kernel
  const bool prefer_single = (tl_policy.jacobi_single == 1 || bt.n > 1) && q <= JS_QMAX;
  if (!prefer_single && q % 4 == 0 && q >= 64 && aligned16(work) && aligned16(evecs)) {
    // cluster kernel: `work` holds q*q floats of column storage followed by q floats of norms and 2 counters
    float* lam_buf = work + (size_t)q * q;
    int* counters = reinterpret_cast<int*>(lam_buf + q);
    CB_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(int), st));
    if (q <= 256) jacobi_cluster_kernel<8, 8><<<JC_CTAS, 256, 0, st>>>(Lc, q, evals, evecs, work, sweeps, counters, lam_buf, 30, kJacobiTol);
    else jacobi_cluster_kernel<16, 8><<<JC_CTAS, 512, 0, st>>>(Lc, q, evals, evecs, work, sweeps, counters, lam_buf, 30, kJacobiTol);
    CB_CHECK_LAUNCH();
    return CB_OK;
  }
  if (q <= JS_QMAX && q % 4 == 0) {
    const size_t smem = ((size_t)q * (q + 4) + 2 * (size_t)q) * sizeof(float);
    static PerDeviceOnce once;
    CB_TRY(opt_in_dynamic_smem(jacobi_smem_kernel, (int)(((size_t)JS_QMAX * (JS_QMAX + 4) + 2 * JS_QMAX) * sizeof(float)), once));
    jacobi_smem_kernel<<<bt.n, JS_THREADS, smem, st>>>(Lc, q, evals, evecs, sweeps, 30, kJacobiTol, bt.stride);
    CB_CHECK_LAUNCH();
    return CB_OK;
  }
  jacobi_kernel<<<1, 1024, 0, st>>>(Lc, q, evals, evecs, work, sweeps, 30, kJacobiTol);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

// ---------------------------------------------------------------- smallest eigenvalue of a dense symmetric matrix
// alg.py:57-64 shifts a Hessian whose smallest eigenvalue is below sigma_reg: H += (sigma_reg - lambda_min) I.  The
// reference reads lambda_min off a full eigendecomposition; here it comes from a Lanczos run with full
// reorthogonalisation (k <= 96 steps, each one symmetric matrix-vector product over all SMs plus a single-CTA
// orthogonalisation step), followed by bisection on the k x k tridiagonal matrix.  The smallest Ritz value converges
// to lambda_min from above: geometrically when lambda_min is separated from the rest of the spectrum (the rank-deficient
// Hessians sigma_reg exists for: lambda_min = 0 with the non-zero eigenvalues bounded away), like range / k^2 otherwise.
constexpr int kLanczosMax = 96;

__global__ void __launch_bounds__(256) symv_kernel(const float* __restrict__ A, const float* __restrict__ v,
                                                    float* __restrict__ y, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* row = A + (size_t)warp * n;
  float acc = 0.f;
  for (int j = lane; j < n; j += 32) acc = fmaf(row[j], v[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[warp] = acc;
}

// work layout: V[(kLanczosMax + 1) x n] | y[n] | alpha[kLanczosMax] | beta[kLanczosMax] | kdone (int) | shift (float)
__global__ void __launch_bounds__(1024) lanczos_init_kernel(float* __restrict__ V, int n, int* __restrict__ kdone) {
  __shared__ double red[32];
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float g = gaussian_from(0x1A2B3C4Dull, (uint64_t)i);
    V[i] = g;
    ss += (double)g * (double)g;
  }
  ss = block_sum(ss, red);
  __shared__ float s_inv;
  if (threadIdx.x == 0) { s_inv = (float)(1.0 / sqrt(ss)); kdone[0] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) V[i] *= s_inv;
}

__global__ void __launch_bounds__(1024)
lanczos_step_kernel(float* __restrict__ V, float* __restrict__ y, float* __restrict__ alpha, float* __restrict__ beta,
                    int* __restrict__ kdone, int j, int n) {
  if (kdone[0] != 0) return;                       // an invariant subspace was found at an earlier step
  __shared__ double red[32];
  __shared__ float s_c;
  const float* vj = V + (size_t)j * n;
  // w = y - sum_i <V_i, y> V_i over the whole basis, twice (classical "twice is enough" reorthogonalisation); the
  // coefficient along V_j of the first pass is alpha_j
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i <= j; ++i) {
      const float* vi = V + (size_t)i * n;
      double d = 0.0;
      for (int e = threadIdx.x; e < n; e += blockDim.x) d += (double)vi[e] * (double)y[e];
      d = block_sum(d, red);
      if (threadIdx.x == 0) { s_c = (float)d; if (pass == 0 && i == j) alpha[j] = (float)d; }
      __syncthreads();
      const float c = s_c;
      for (int e = threadIdx.x; e < n; e += blockDim.x) y[e] = fmaf(-c, vi[e], y[e]);
      __syncthreads();
    }
  }
  (void)vj;
  double ss = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) ss += (double)y[e] * (double)y[e];
  ss = block_sum(ss, red);
  __shared__ float s_beta;
  if (threadIdx.x == 0) {
    s_beta = (float)sqrt(ss);
    beta[j] = s_beta;
    // breakdown: the Krylov space is invariant, the tridiagonal matrix of order j + 1 already holds exact eigenvalues
    if (!(s_beta > 1e-6f * fmaxf(fabsf(alpha[j]), 1e-30f))) kdone[0] = j + 1;
  }
  __syncthreads();
  if (kdone[0] != 0) return;
  const float inv = 1.f / s_beta;
  float* vn = V + (size_t)(j + 1) * n;
  for (int e = threadIdx.x; e < n; e += blockDim.x) vn[e] = y[e] * inv;
}

// smallest eigenvalue of the k x k tridiagonal matrix (alpha, beta) by bisection on the Sturm count; one thread
__global__ void lanczos_finish_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, const int* __restrict__ kdone,
                                      int kmax, float sigma_reg, float* __restrict__ out /* [0] shift, [1] lambda_min, [2] lambda_max */) {
  const int k = kdone[0] != 0 ? kdone[0] : kmax;
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < k; ++i) {
    const double off = (i > 0 ? fabs((double)beta[i - 1]) : 0.0) + (i + 1 < k ? fabs((double)beta[i]) : 0.0);
    lo = fmin(lo, (double)alpha[i] - off);
    hi = fmax(hi, (double)alpha[i] + off);
  }
  // count of eigenvalues < x
  auto count_below = [&](double x) {
    int c = 0;
    double d = 1.0;
    for (int i = 0; i < k; ++i) {
      const double b2 = i > 0 ? (double)beta[i - 1] * (double)beta[i - 1] : 0.0;
      d = ((double)alpha[i] - x) - (i > 0 ? b2 / d : 0.0);
      if (d == 0.0) d = -1e-300;
      if (d < 0.0) ++c;
    }
    return c;
  };
  for (int it = 0; it < 200 && hi - lo > 1e-14 * fmax(fabs(lo), fabs(hi)) + 1e-300; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (count_below(mid) >= 1) hi = mid; else lo = mid;
  }
  const float lam = (float)(0.5 * (lo + hi));
  out[1] = lam;
  out[0] = lam < sigma_reg ? sigma_reg - lam : 0.f;      // alg.py:59-63
  // largest Ritz value (step size of the dense-Hessian proximal-gradient loop)
  double lo2 = lo, hi2 = -1e300;
  for (int i = 0; i < k; ++i) {
    const double off = (i > 0 ? fabs((double)beta[i - 1]) : 0.0) + (i + 1 < k ? fabs((double)beta[i]) : 0.0);
    hi2 = fmax(hi2, (double)alpha[i] + off);
  }
  for (int it = 0; it < 200 && hi2 - lo2 > 1e-14 * fmax(fabs(lo2), fabs(hi2)) + 1e-300; ++it) {
    const double mid = 0.5 * (lo2 + hi2);
    if (count_below(mid) >= k) hi2 = mid; else lo2 = mid;
  }
  out[2] = (float)(0.5 * (lo2 + hi2));
}

__global__ void __launch_bounds__(256) add_diag_kernel(float* __restrict__ A, int n, const float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float s = shift[0];
  if (i < n && s != 0.f) A[(size_t)i * n + i] += s;
}

size_t min_eig_shift_workspace_floats(int64_t n) {
  const int64_t k = n < kLanczosMax ? n : kLanczosMax;
  return (size_t)((k + 2) * n + 2 * kLanczosMax + 8);
}

// Hs (n x n symmetric, in place) += max(0, sigma_reg - lambda_min(Hs)) I; stats (optional, 2 floats): shift, lambda_min
int min_eig_shift(float* Hs, int64_t n, float sigma_reg, float* work, float* stats, cudaStream_t st, int nstats) {
  if (Hs == nullptr || work == nullptr || n <= 0) return CB_ERR_ARG;
  const int k = (int)(n < kLanczosMax ? n : kLanczosMax);
  float* V = work;
  float* y = V + (size_t)(k + 1) * n;
  float* alpha = y + n;
  float* beta = alpha + kLanczosMax;
  int* kdone = reinterpret_cast<int*>(beta + kLanczosMax);
  float* out = reinterpret_cast<float*>(kdone + 2);
  lanczos_init_kernel<<<1, 1024, 0, st>>>(V, (int)n, kdone);
  CB_CHECK_LAUNCH();
  for (int j = 0; j < k; ++j) {
    symv_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(Hs, V + (size_t)j * n, y, (int)n);
    CB_CHECK_LAUNCH();
    lanczos_step_kernel<<<1, 1024, 0, st>>>(V, y, alpha, beta, kdone, j, (int)n);
    CB_CHECK_LAUNCH();
  }
  lanczos_finish_kernel<<<1, 1, 0, st>>>(alpha, beta, kdone, k, sigma_reg, out);
  CB_CHECK_LAUNCH();
  add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Hs, (int)n, out);
  CB_CHECK_LAUNCH();
  if (stats != nullptr && nstats > 0)
    CB_CUDA(cudaMemcpyAsync(stats, out, (size_t)(nstats < 3 ? nstats : 3) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return CB_OK;
}

}  // namespace cb

#ifdef CB_MEASURE
extern "C" void cb_set_chol_timing(void* stamps_dev) { cb::g_chol_timing = reinterpret_cast<long long*>(stamps_dev); }
#endif

extern "C" int cb_cholesky_inverse_f32(float* G, int64_t q, float* Linv, int* status, void* stream) {
  return cb::cholesky_inverse(G, (int)q, Linv, status, (cudaStream_t)stream);
}

extern "C" size_t cb_min_eig_shift_workspace_bytes(int64_t n) {
  return n > 0 ? cb::min_eig_shift_workspace_floats(n) * sizeof(float) : 0;
}
extern "C" int cb_min_eig_shift_f32(float* H, int64_t n, float sigma_reg, float* stats, void* ws, size_t ws_bytes, void* stream) {
  if (H == nullptr || ws == nullptr || n <= 0) return CB_ERR_ARG;
  if (ws_bytes < cb_min_eig_shift_workspace_bytes(n)) return CB_ERR_WORKSPACE;
  return cb::min_eig_shift(H, n, sigma_reg, reinterpret_cast<float*>(ws), stats, (cudaStream_t)stream);
}

extern "C" int cb_convex_dense_prepare(const float* H, int64_t n, float floor, float* Hs, float* stats3, void* ws,
                                       size_t ws_bytes, void* stream) {
  if (H == nullptr || Hs == nullptr || stats3 == nullptr || ws == nullptr || n <= 0) return CB_ERR_ARG;
  if (ws_bytes < cb_min_eig_shift_workspace_bytes(n)) return CB_ERR_WORKSPACE;
  CB_TRY(cb::symmetrize(H, n, Hs, (cudaStream_t)stream));
  return cb::min_eig_shift(Hs, n, floor, reinterpret_cast<float*>(ws), stats3, (cudaStream_t)stream, 3);
}

extern "C" int cb_jacobi_eigh_from_chol_f32(const float* Lc, int64_t q, float* evals, float* evecs, float* work,
                                            int* sweeps, void* stream) {
  return cb::jacobi_eigh_from_chol(Lc, (int)q, evals, evecs, work, sweeps, (cudaStream_t)stream);
}

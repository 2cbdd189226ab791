// Small dense factorizations that sit between the big contractions:
//   * blocked Cholesky + triangular inverse of a q x q Gram matrix (q <= 512), used by the
//     CholeskyQR re-orthonormalisation of the sketch and by the normal-equation solves of
//     the LPLR updates (replaces torch.linalg.lstsq's QR, alg.py:163,175);
//   * one-sided Jacobi eigensolver on the Cholesky factor, the Rayleigh-Ritz step that
//     replaces the truncated SVD of LR_init (alg.py:214-217).
// Both are single-CTA, latency-bound kernels: the matrices are at most 1 MiB and stay in
// L1/L2; panels are staged in shared memory.
#include "common.cuh"
#include "internal.h"

namespace cb {

constexpr int NB = 32;          // panel width
constexpr int QMAX = 512;       // largest supported matrix
constexpr int TMAXR = QMAX - NB;  // most rows below a diagonal block

struct CholSmem {
  float D[NB][NB + 1];             // diagonal block / its Cholesky factor
  float Di[NB][NB + 1];            // inverse of the factor
  float Praw[TMAXR][NB + 1];       // panel rows before the triangular solve (row-major)
  float Pt[NB][TMAXR + 4];         // solved panel, transposed: Pt[c][row]
  float diag0[QMAX];               // backup of the diagonal for the ridge retry
  int fail;
  float mean_diag;
};

// In-place Cholesky of the lower triangle of G (row-major, leading dimension q).  The strict
// upper triangle is never written, so together with diag0 it is a backup of the input.
__device__ void chol_factor(float* __restrict__ G, int q, float* __restrict__ Linv, CholSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < q; k0 += NB) {
    const int nb = min(NB, q - k0);
    // 1. diagonal block -> shared
    {
      const int i = tid >> 5, j = tid & 31;
      float v = 0.f;
      if (i < nb && j <= i) v = G[(size_t)(k0 + i) * q + k0 + j];
      s.D[i][j] = v;
      s.Di[i][j] = 0.f;
    }
    __syncthreads();
    // 2. warp 0 factorises it (lane = row) and inverts the factor (lane = column)
    if (warp == 0) {
      for (int j = 0; j < nb; ++j) {
        float d = s.D[j][j];
        if (!(d > 0.f) || !isfinite(d)) { if (lane == 0) s.fail = 1; d = 1.f; }
        const float sd = sqrtf(d), inv = 1.f / sd;
        float lij = 0.f;
        __syncwarp();
        if (lane == j) s.D[j][j] = sd;
        if (lane > j && lane < nb) { lij = s.D[lane][j] * inv; s.D[lane][j] = lij; }
        __syncwarp();
        if (lane > j && lane < nb)
          for (int c = j + 1; c <= lane; ++c) s.D[lane][c] -= lij * s.D[c][j];
        __syncwarp();
      }
      // inverse: column `lane` by forward substitution, loops kept warp-uniform
      if (lane < nb) s.Di[lane][lane] = 1.f / s.D[lane][lane];
      __syncwarp();
      for (int i = 1; i < nb; ++i) {
        float acc = 0.f;
        for (int k = 0; k < i; ++k)
          if (k >= lane) acc += s.D[i][k] * s.Di[k][lane];
        if (lane < i) s.Di[i][lane] = -acc / s.D[i][i];
        __syncwarp();
      }
    }
    __syncthreads();
    // write the factor back, publish the inverse block
    {
      const int i = tid >> 5, j = tid & 31;
      if (i < nb && j <= i) {
        G[(size_t)(k0 + i) * q + k0 + j] = s.D[i][j];
        if (Linv != nullptr) Linv[(size_t)(k0 + i) * q + k0 + j] = s.Di[i][j];
      }
    }
    const int T = q - k0 - nb;  // rows below the block
    if (T <= 0) { __syncthreads(); continue; }
    // 3a. stage the raw panel (coalesced along the row)
    for (int e = tid; e < T * NB; e += blockDim.x) {
      const int i = e >> 5, k = e & 31;
      s.Praw[i][k] = (k < nb) ? G[(size_t)(k0 + nb + i) * q + k0 + k] : 0.f;
    }
    __syncthreads();
    // 3b. P = Praw * Lkk^-T : P[i][c] = sum_{k<=c} Praw[i][k] * Di[c][k]; lanes = consecutive rows
    for (int e = tid; e < T * nb; e += blockDim.x) {
      const int c = e / T, i = e - c * T;
      float acc = 0.f;
      for (int k = 0; k <= c; ++k) acc += s.Praw[i][k] * s.Di[c][k];
      s.Pt[c][i] = acc;
      G[(size_t)(k0 + nb + i) * q + k0 + c] = acc;
    }
    // zero-pad the panel to a multiple of 4 rows so 128-bit reads below stay in bounds
    for (int e = tid; e < 4 * NB; e += blockDim.x) {
      const int c = e >> 2, i = T + (e & 3);
      if (i < TMAXR + 4) s.Pt[c][i] = 0.f;
    }
    if (nb < NB)
      for (int e = tid; e < (NB - nb) * (TMAXR + 4); e += blockDim.x) s.Pt[nb + e / (TMAXR + 4)][e % (TMAXR + 4)] = 0.f;
    __syncthreads();
    // 4. trailing update on the lower triangle, 4x4 register tiles
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
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ti * 4 + a;
        if (i >= T) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = tj * 4 + b;
          if (j > i) continue;
          G[(size_t)(k0 + nb + i) * q + k0 + nb + j] -= acc[a][b];
        }
      }
    }
    __syncthreads();
  }
}

// Linv = Lc^-1 by block forward substitution; diagonal blocks already hold their inverses.
__device__ void tri_inverse(const float* G, int q, float* Linv, CholSmem& s) {
  const int tid = threadIdx.x;
  const int nblk = (q + NB - 1) / NB;
  // strict upper triangle and not-yet-computed blocks start at zero
  for (int e = tid; e < q * q; e += blockDim.x) {
    const int i = e / q, j = e - i * q;
    if ((i / NB) != (j / NB) || j > i) Linv[e] = 0.f;
  }
  __syncthreads();
  for (int bi = 1; bi < nblk; ++bi) {
    const int r0 = bi * NB, nbi = min(NB, q - r0), W = r0;  // W columns to the left
    // stage L[bi, 0:W] as Lrow[i][k] in Pt (NB x (TMAXR+4)) and Di_bi in Di
    for (int e = tid; e < nbi * W; e += blockDim.x) {
      const int i = e / W, k = e - i * W;
      s.Pt[i][k] = G[(size_t)(r0 + i) * q + k];
    }
    {
      const int i = tid >> 5, j = tid & 31;
      s.Di[i][j] = (i < nbi && j <= i) ? Linv[(size_t)(r0 + i) * q + r0 + j] : 0.f;
    }
    __syncthreads();
    // S[i][col] = sum_{k = blockstart(col)}^{W-1} Lrow[i][k] * Linv[k][col]; staged in Praw (as [col][i])
    for (int e = tid; e < nbi * W; e += blockDim.x) {
      const int i = e / W, col = e - i * W;
      float acc = 0.f;
      for (int k = (col / NB) * NB; k < W; ++k) acc = fmaf(s.Pt[i][k], Linv[(size_t)k * q + col], acc);
      s.Praw[col][i] = acc;
    }
    __syncthreads();
    // Linv[bi][col] = -Di_bi * S
    for (int e = tid; e < nbi * W; e += blockDim.x) {
      const int i = e / W, col = e - i * W;
      float acc = 0.f;
      for (int k = 0; k <= i; ++k) acc = fmaf(s.Di[i][k], s.Praw[col][k], acc);
      Linv[(size_t)(r0 + i) * q + col] = -acc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024, 1)
chol_inv_kernel(float* __restrict__ G, int q, float* __restrict__ Linv, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem& s = *reinterpret_cast<CholSmem*>(smem_raw);
  const int tid = threadIdx.x;
  double dsum = 0.0;
  for (int i = tid; i < q; i += blockDim.x) { s.diag0[i] = G[(size_t)i * q + i]; }
  if (tid == 0) s.fail = 0;
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < q; ++i) dsum += (double)s.diag0[i];
    s.mean_diag = (float)(dsum / (double)q);
  }
  __syncthreads();
  int retries = 0;
  const float ridges[3] = {1e-6f, 1e-4f, 1e-2f};
  while (true) {
    chol_factor(G, q, Linv, s);
    __syncthreads();
    const int failed = s.fail;
    __syncthreads();
    if (!failed || retries >= 3) { if (failed) retries = 99; break; }
    // restore the lower triangle from the untouched upper one, add a ridge, retry
    const float ridge = ridges[retries] * fmaxf(s.mean_diag, 1e-30f);
    for (int e = tid; e < q * q; e += blockDim.x) {
      const int i = e / q, j = e - i * q;
      if (j < i) G[e] = G[(size_t)j * q + i];
      else if (j == i) G[e] = s.diag0[i] + ridge;
    }
    if (tid == 0) s.fail = 0;
    ++retries;
    __syncthreads();
  }
  if (Linv != nullptr) tri_inverse(G, q, Linv, s);
  if (tid == 0 && status != nullptr) atomicMax(status, retries);
}

int cholesky_inverse(float* G, int q, float* Linv, int* status, cudaStream_t st) {
  if (G == nullptr || q <= 0) return CB_ERR_ARG;
  if (q > QMAX) return CB_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CholSmem)));
    attr_set = true;
  }
  chol_inv_kernel<<<1, 1024, sizeof(CholSmem), st>>>(G, q, Linv, status);
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
          const float zeta = (be - al) / (2.f * ga);
          const float tt = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
          // IEEE sqrt/divide: rsqrtf's 2-ulp bias accumulates over ~q rotations per vector per sweep
          const float c = __fdiv_rn(1.f, __fsqrt_rn(fmaf(tt, tt, 1.f))), sn = c * tt;
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
    if (rot == 0) { ++sweep; break; }
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

int jacobi_eigh_from_chol(const float* Lc, int q, float* evals, float* evecs, float* work, int* sweeps,
                          cudaStream_t st) {
  if (Lc == nullptr || evals == nullptr || evecs == nullptr || work == nullptr || q <= 0) return CB_ERR_ARG;
  if (q > QMAX) return CB_ERR_UNSUPPORTED;
  jacobi_kernel<<<1, 1024, 0, st>>>(Lc, q, evals, evecs, work, sweeps, 30, 1e-6f);
  CB_CHECK_LAUNCH();
  return CB_OK;
}

}  // namespace cb

extern "C" int cb_cholesky_inverse_f32(float* G, int64_t q, float* Linv, int* status, void* stream) {
  return cb::cholesky_inverse(G, (int)q, Linv, status, (cudaStream_t)stream);
}

extern "C" int cb_jacobi_eigh_from_chol_f32(const float* Lc, int64_t q, float* evals, float* evecs, float* work,
                                            int* sweeps, void* stream) {
  return cb::jacobi_eigh_from_chol(Lc, (int)q, evals, evecs, work, sweeps, (cudaStream_t)stream);
}

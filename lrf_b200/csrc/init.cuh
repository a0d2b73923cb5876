// SVD initialisation, stage three: from the top-R right singular vectors E (FP64, from eig.cuh)
// and sigma to the scaled factors of SVDInit (lrf/factorization/qmf.py:42-52):
//     u0 = U_R * sqrt(s),  v0 = V_R * sqrt(s),   U_R = X V_R / s
// u_hat and v_hat are rounded to f32 once (as LAPACK hands them to torch), s is rounded to f32,
// sqrt and the scaling multiply are f32 ops exactly like torch.sqrt / einsum there.
// Columns beyond min(M, N) and columns with sigma == 0 (rank-deficient input) are zero.
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {

constexpr int kProjRows = 128;  // rows per CTA tile (= threads per CTA)

// grid = (row tiles, matrices), 128 threads.  The X tile is staged through shared memory with
// coalesced loads (row stride N+1: conflict-free for the thread-per-row reads that follow).
__global__ void __launch_bounds__(kProjRows)
svd_project_kernel(const float* __restrict__ X, long long x_stride, int M, int N, int R,
                   const double* __restrict__ evec, const double* __restrict__ sigma,
                   float* __restrict__ U0, float* __restrict__ V0) {
  LRFB_DYN_SMEM(smem_raw);
  double* ev = reinterpret_cast<double*>(smem_raw);               // [N][R]
  float* xt = reinterpret_cast<float*>(ev + (size_t)N * R);       // [kProjRows][N+1]
  const int mat = blockIdx.y;
  const int keep = min(R, min(M, N));
  const int XS = N + 1;
  for (int i = threadIdx.x; i < N * R; i += blockDim.x) ev[i] = evec[(size_t)mat * N * R + i];
  const float* x = X + (size_t)mat * x_stride;
  float* u0 = U0 + (size_t)mat * M * R;
  float* v0 = V0 + (size_t)mat * N * R;
  const double* sg = sigma + (size_t)mat * R;
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < N * R; i += blockDim.x) {
      int r = i % R;
      float val = 0.0f;
      if (r < keep && sg[r] > 0.0) val = __fmul_rn((float)ev[i], __fsqrt_rn((float)sg[r]));
      v0[i] = val;
    }
  }
  for (int r0 = blockIdx.x * kProjRows; r0 < M; r0 += gridDim.x * kProjRows) {
    const int valid = min(kProjRows, M - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < valid * N; e += blockDim.x) {
      int r = e / N, c = e - r * N;
      xt[r * XS + c] = x[(size_t)r0 * N + e];
    }
    __syncthreads();
    const int m = threadIdx.x;
    if (m < valid) {
      for (int c0 = 0; c0 < R; c0 += 4) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < N; ++k) {
          const double xv = (double)xt[m * XS + k];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j < R) acc[j] = fma(xv, ev[k * R + c0 + j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = c0 + j;
          if (r >= R) break;
          float val = 0.0f;
          if (r < keep && sg[r] > 0.0) val = __fmul_rn((float)(acc[j] / sg[r]), __fsqrt_rn((float)sg[r]));
          u0[(size_t)(r0 + m) * R + r] = val;
        }
      }
    }
  }
}

}  // namespace lrfb

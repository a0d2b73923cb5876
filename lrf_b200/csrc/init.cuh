// SVD initialisation, stage three: from the top-R right singular vectors E (FP64, from eig.cuh)
// and sigma to the scaled factors of SVDInit (lrf/factorization/qmf.py:42-52):
//     u0 = U_R * sqrt(s),  v0 = V_R * sqrt(s),   U_R = X V_R / s
// u_hat and v_hat are rounded to f32 once (as LAPACK hands them to torch), s is rounded to f32,
// sqrt and the scaling multiply are f32 ops exactly like torch.sqrt / einsum there.
// Columns beyond min(M, N) and columns with sigma == 0 (rank-deficient input) are zero.
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {

__global__ void __launch_bounds__(128)
svd_project_kernel(const float* __restrict__ X, long long x_stride, int M, int N, int R,
                   const double* __restrict__ evec, const double* __restrict__ sigma,
                   float* __restrict__ U0, float* __restrict__ V0) {
  LRFB_DYN_SMEM(smem_raw);
  double* ev = reinterpret_cast<double*>(smem_raw);  // [N][R]
  const int mat = blockIdx.y;
  const int keep = min(R, min(M, N));
  for (int i = threadIdx.x; i < N * R; i += blockDim.x) ev[i] = evec[(size_t)mat * N * R + i];
  __syncthreads();
  const float* x = X + (size_t)mat * x_stride;
  float* u0 = U0 + (size_t)mat * M * R;
  float* v0 = V0 + (size_t)mat * N * R;
  const double* sg = sigma + (size_t)mat * R;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < N * R; i += blockDim.x) {
      int r = i % R;
      float val = 0.0f;
      if (r < keep && sg[r] > 0.0) val = __fmul_rn((float)ev[i], __fsqrt_rn((float)sg[r]));
      v0[i] = val;
    }
  }
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
    for (int r0 = 0; r0 < R; r0 += 4) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < N; ++k) {
        double xv = (double)x[(size_t)m * N + k];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (r0 + j < R) acc[j] = fma(xv, ev[k * R + r0 + j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r = r0 + j;
        if (r >= R) break;
        float val = 0.0f;
        if (r < keep && sg[r] > 0.0) val = __fmul_rn((float)(acc[j] / sg[r]), __fsqrt_rn((float)sg[r]));
        u0[(size_t)m * R + r] = val;
      }
    }
  }
}

}  // namespace lrfb

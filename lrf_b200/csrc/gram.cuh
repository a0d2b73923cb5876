// FP64 Gram matrix G = X^T X of a batch of (M x N) f32 patch matrices — the first stage of the SVD
// initialisation that replaces torch.linalg.svd (lrf/factorization/qmf.py:44).  SURVEY H2: an FP32
// Gram is not accurate enough to land on LAPACK's singular vectors; products of two f32 are exact in
// f64 and the accumulation error is ~1e-16 relative, so the result is order-independent in practice.
//
// Mapping: the upper triangle of G is cut into 4x4 blocks; each thread owns up to BPT blocks
// (16 DFMA per block per row), rows are staged through shared memory already converted to f64
// (one F2F per element instead of one per use).  grid = (row splits, matrices); a split writes a
// partial Gram which gram_reduce_kernel sums in fixed order; grid.z cuts the block list when one CTA's
// registers cannot hold all of it (N = 256).
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {

constexpr int kGramTileRows = 16;

template <int BPT>
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ X, long long x_stride, int M, int N,
                            double* __restrict__ Gout, int n_split) {
  LRFB_DYN_SMEM(smem_raw);
  double* tile = reinterpret_cast<double*>(smem_raw);  // [kGramTileRows][Npad]
  const int Npad = (N + 3) & ~3;
  const int nb = Npad / 4;
  const int nblocks = nb * (nb + 1) / 2;
  const int mat = blockIdx.y, split = blockIdx.x;
  const float* x = X + (size_t)mat * x_stride;

  int bi[BPT], bj[BPT];
  double acc[BPT][16];
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    int idx = blockIdx.z * (blockDim.x * BPT) + threadIdx.x + b * blockDim.x;
    bi[b] = -1, bj[b] = 0;
    if (idx < nblocks) {  // idx -> (i, j), i <= j, row-major over the upper triangle
      int i = 0, rem = idx;
      while (rem >= nb - i) rem -= nb - i, ++i;
      bi[b] = i, bj[b] = i + rem;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[b][k] = 0.0;
  }

  const int n_tiles = (M + kGramTileRows - 1) / kGramTileRows;
  for (int t = split; t < n_tiles; t += n_split) {
    const int r0 = t * kGramTileRows;
    __syncthreads();
    for (int e = threadIdx.x; e < kGramTileRows * Npad; e += blockDim.x) {
      int r = e / Npad, c = e - r * Npad;
      double v = 0.0;
      if (r0 + r < M && c < N) v = (double)x[(size_t)(r0 + r) * N + c];
      tile[e] = v;
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      if (bi[b] < 0) continue;
      const double* pa = tile + 4 * bi[b];
      const double* pb = tile + 4 * bj[b];
#pragma unroll 4
      for (int r = 0; r < kGramTileRows; ++r) {
        double a[4], c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = pa[r * Npad + k], c[k] = pb[r * Npad + k];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[b][i * 4 + j] = fma(a[i], c[j], acc[b][i * 4 + j]);
      }
    }
  }

  double* g = Gout + ((size_t)mat * n_split + split) * (size_t)N * N;
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    if (bi[b] < 0) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gi = 4 * bi[b] + i, gj = 4 * bj[b] + j;
        if (gi < N && gj < N) {
          g[(size_t)gi * N + gj] = acc[b][i * 4 + j];
          g[(size_t)gj * N + gi] = acc[b][i * 4 + j];
        }
      }
  }
}

// G[mat] = sum over splits (ascending) of partial[mat][split]
__global__ void gram_reduce_kernel(const double* __restrict__ partial, double* __restrict__ G, int nn,
                                   int n_split) {
  const int mat = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < n_split; ++k) s += partial[((size_t)mat * n_split + k) * nn + e];
    G[(size_t)mat * nn + e] = s;
  }
}

}  // namespace lrfb

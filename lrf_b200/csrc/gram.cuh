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


// ---------------------------------------------------------------------------------------------------
// N = 64 on the FP64 tensor-core path: mma.sync.m8n8k4.f64 (DMMA).  The upper triangle of G is 36
// blocks of 8x8; warp w of a 4-warp CTA owns block rows w and 7-w (9 blocks), so every warp issues 9
// DMMAs per 4 rows of X from at most 10 fragment loads.  Rows are staged as f64 with a row stride of
// 68 doubles (fragment loads hit 32 distinct banks per half warp).
// ---------------------------------------------------------------------------------------------------
constexpr int kDmmaTileRows = 32;
constexpr int kDmmaLd = 68;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
#ifdef LRFB_SIM
  const int lane = threadIdx.x & 31;
  const int i = lane >> 2, j0 = (lane & 3) * 2;
  for (int k = 0; k < 4; ++k) {
    double ak = __shfl_sync(0xffffffffu, a, i * 4 + k);
    double b0 = __shfl_sync(0xffffffffu, b, j0 * 4 + k);
    double b1 = __shfl_sync(0xffffffffu, b, (j0 + 1) * 4 + k);
    c0 = fma(ak, b0, c0);
    c1 = fma(ak, b1, c1);
  }
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
#endif
}

// the 9 blocks of warp W for one tile: block rows W and 7-W (all indices compile-time)
template <int W>
__device__ __forceinline__ void gram64_tile_mma(const double* __restrict__ tile, double (&acc)[9][2], int lane) {
  constexpr int RA = W, RB = 7 - W;
#pragma unroll 2
  for (int m0 = 0; m0 < kDmmaTileRows; m0 += 4) {
    const double* base = tile + (m0 + (lane & 3)) * kDmmaLd + (lane >> 2);
    double f[8];
#pragma unroll
    for (int b = RA; b < 8; ++b) f[b] = base[8 * b];  // fragments of block columns RA..7 (RB >= RA)
#pragma unroll
    for (int jb = RA; jb < 8; ++jb) dmma_m8n8k4(acc[jb - RA][0], acc[jb - RA][1], f[RA], f[jb]);
#pragma unroll
    for (int jb = RB; jb < 8; ++jb)
      dmma_m8n8k4(acc[(8 - RA) + jb - RB][0], acc[(8 - RA) + jb - RB][1], f[RB], f[jb]);
  }
}

__global__ void __launch_bounds__(128)
gram64_dmma_kernel(const float* __restrict__ X, long long x_stride, int M, double* __restrict__ Gout,
                   int n_split) {
  __align__(16) __shared__ double tile[kDmmaTileRows * kDmmaLd];
  const int mat = blockIdx.y, split = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x = X + (size_t)mat * x_stride;
  double acc[9][2];
#pragma unroll
  for (int b = 0; b < 9; ++b) acc[b][0] = acc[b][1] = 0.0;

  const int n_tiles = (M + kDmmaTileRows - 1) / kDmmaTileRows;
  for (int t = split; t < n_tiles; t += n_split) {
    const int r0 = t * kDmmaTileRows;
    __syncthreads();
    for (int e = threadIdx.x; e < kDmmaTileRows * 16; e += 128) {  // one float4 per item
      const int r = e >> 4, c4 = (e & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < M) v = *reinterpret_cast<const float4*>(x + (size_t)(r0 + r) * 64 + c4);
      double* dst = tile + r * kDmmaLd + c4;
      dst[0] = (double)v.x, dst[1] = (double)v.y, dst[2] = (double)v.z, dst[3] = (double)v.w;
    }
    __syncthreads();
    switch (warp) {
      case 0: gram64_tile_mma<0>(tile, acc, lane); break;
      case 1: gram64_tile_mma<1>(tile, acc, lane); break;
      case 2: gram64_tile_mma<2>(tile, acc, lane); break;
      default: gram64_tile_mma<3>(tile, acc, lane); break;
    }
  }

  // blocks 0..(7-warp): block row `warp`, columns warp..7; then block row 7-warp, columns (7-warp)..7
  double* g = Gout + ((size_t)mat * n_split + split) * 4096;
#pragma unroll
  for (int n = 0; n < 9; ++n) {
    const int first = 8 - warp;  // number of blocks in block row `warp`
    const int ib = n < first ? warp : 7 - warp;
    const int jb = n < first ? warp + n : (7 - warp) + (n - first);
    const int gi = 8 * ib + (lane >> 2), gj = 8 * jb + (lane & 3) * 2;
    g[gi * 64 + gj] = acc[n][0], g[gi * 64 + gj + 1] = acc[n][1];
    g[gj * 64 + gi] = acc[n][0], g[(gj + 1) * 64 + gi] = acc[n][1];
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

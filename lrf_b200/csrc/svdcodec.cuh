// Kernels of the SVD baseline codec (lrf.svd_encode / lrf.svd_decode, RGB + patch branch,
// lrf/compression/svd.py:156-193, :297-361), beyond the shared front end / Gram / eigen-solver:
//   svd_project_kernel   u = U_R * sqrt(s) with U_R = X V_R / s            (svd.py:179-183)
//   minmax / quantize    quantize(t, uint8): scale = (max-min)/255, q = trunc(clamp((t-min)/scale, 0, 255))
//                                                                           (compression/utils.py:185-220)
//   svd_decode_kernel    dequantize (:223-243), u @ v.T (ascending-r FMA chain), depatchify, unpad,
//                        clamp + truncate to uint8                           (svd.py:316-359)
// Unlike the QMF path the float factors themselves are the payload, so U is materialised here.
#pragma once
#include "decode.cuh"

namespace lrfb {

constexpr int kProjRows = 128;  // rows per CTA tile
constexpr int kProjCols = 16;   // columns of U per pass

// grid = (row tiles, matrices), 128 threads.  u[m][r] = f32(X[m]·E[:,r] / sigma_r) * sqrtf(f32(sigma_r)), the dot
// product in f64 with k ascending (u_hat is then the correctly rounded f32 of the true left singular vector entry).
// Register tile: warp w owns columns 4w..4w+3 of a 16-column pass, lane l the rows l, l+32, l+64, l+96 of the tile —
// 16 independent DFMA chains per thread fed by 4 row loads (16 bytes each when N % 4 == 0; the other 16 bytes of the
// sector are the next step's, an L1 hit) and 2 broadcast 16-byte reads of E per k, so the FP64 pipe is the limit.
template <bool VEC>
__global__ void __launch_bounds__(kProjRows)
svd_project_kernel(const float* __restrict__ X, long long x_stride, int M, int N, int R,
                   const double* __restrict__ evec, const double* __restrict__ sigma, float* __restrict__ U) {
  LRFB_DYN_SMEM(smem_raw);
  double* ev = reinterpret_cast<double*>(smem_raw);  // [N][16]: the current pass's columns, zero padded
  const int mat = blockIdx.y;
  const int keep = min(R, min(M, N));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x = X + (size_t)mat * x_stride;
  float* u = U + (size_t)mat * M * R;
  const double* sg = sigma + (size_t)mat * R;
  const double* E = evec + (size_t)mat * N * R;
  for (int c0 = 0; c0 < R; c0 += kProjCols) {
    __syncthreads();
    for (int i = threadIdx.x; i < N * kProjCols; i += blockDim.x) {
      const int k = i / kProjCols, j = i % kProjCols;
      ev[i] = c0 + j < R ? E[(size_t)k * R + c0 + j] : 0.0;
    }
    __syncthreads();
    for (int r0 = blockIdx.x * kProjRows; r0 < M; r0 += gridDim.x * kProjRows) {
      const float* xr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[i] = x + (size_t)min(r0 + lane + 32 * i, M - 1) * N;  // rows past M: clamped, not stored
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      const double* ew = ev + 4 * warp;
      if (VEC) {
        for (int k = 0; k < N; k += 4) {
          float4 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xr[i] + k);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            double e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = ew[(k + kk) * kProjCols + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const double xd = (double)(kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w);
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fma(xd, e[j], acc[i][j]);
            }
          }
        }
      } else {
        for (int k = 0; k < N; ++k) {
          double e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = ew[k * kProjCols + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double xd = (double)xr[i][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(xd, e[j], acc[i][j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = r0 + lane + 32 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = c0 + 4 * warp + j;
          if (m < M && r < R) {
            float val = 0.0f;
            if (r < keep && sg[r] > 0.0) val = __fmul_rn((float)(acc[i][j] / sg[r]), __fsqrt_rn((float)sg[r]));
            u[(size_t)m * R + r] = val;
          }
        }
      }
    }
  }
}

// per-matrix min and max of a float array; grid.y = matrices, one block each (n is small: <= M*R)
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ t, long long per_mat, float* __restrict__ out /* [n][2] */) {
  __shared__ float smin[8], smax[8];
  const float* p = t + (size_t)blockIdx.x * per_mat;
  float lo = p[0], hi = p[0];
  for (long long i = threadIdx.x; i < per_mat; i += blockDim.x) lo = fminf(lo, p[i]), hi = fmaxf(hi, p[i]);
  for (int o = 16; o; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = lo, smax[threadIdx.x >> 5] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) lo = fminf(lo, smin[w]), hi = fmaxf(hi, smax[w]);
    out[2 * blockIdx.x] = lo, out[2 * blockIdx.x + 1] = hi;
  }
}

// q = uint8(clamp((t - min)/scale + 0, 0, 255)), scale = (max - min)/255, written fiber-major [R][rows];
// qparams[mat] = {scale, min}.  Every op separately rounded as in the reference's torch expression.
__global__ void __launch_bounds__(256)
quantize_u8_kernel(const float* __restrict__ t, int rows, int R, const float* __restrict__ mm,
                   unsigned char* __restrict__ codes, long long code_stride, float* __restrict__ qparams,
                   long long qp_stride) {
  const int mat = blockIdx.y;
  const float lo = mm[2 * mat], hi = mm[2 * mat + 1];
  const float scale = __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
  if (blockIdx.x == 0 && threadIdx.x == 0) qparams[(size_t)mat * qp_stride] = scale, qparams[(size_t)mat * qp_stride + 1] = lo;
  const float* p = t + (size_t)mat * rows * R;
  unsigned char* q = codes + (size_t)mat * code_stride;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < rows * R; e += gridDim.x * blockDim.x) {
    const int r = e / rows, m = e - r * rows;  // output order: fiber-major
    float v = __fadd_rn(__fdiv_rn(__fsub_rn(p[(size_t)m * R + r], lo), scale), 0.0f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);  // NaN (0/0 for a constant tensor) -> fmaxf picks 0, as torch.clamp would give NaN -> uint8 0
    q[e] = (unsigned char)(int)v;
  }
}

struct SvdDecodeParams {
  int H, W, p, q, n_img, R;
  PlaneGeom g;
  long long record_bytes, u_off, v_off;
};

// one thread per output pixel and channel triple; qparams[img] = {scale_u, min_u, scale_v, min_v, qmin_u, qmin_v}
__global__ void __launch_bounds__(256)
svd_decode_kernel(const unsigned char* __restrict__ codes, const float* __restrict__ qparams,
                  unsigned char* __restrict__ out, SvdDecodeParams P) {
  const size_t hw = (size_t)P.H * P.W;
  const int ncols = 3 * P.p * P.q;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const unsigned char* rec = codes + (size_t)im * P.record_bytes;
    const float* qp = qparams + (size_t)im * 6;
    const float su = qp[0], mu = qp[1], sv = qp[2], mv = qp[3], qu0 = qp[4], qv0 = qp[5];
    unsigned char* o = out + (size_t)im * 3 * hw;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < hw; e += (size_t)gridDim.x * blockDim.x) {
      const int y = (int)(e / P.W), x = (int)(e - (size_t)y * P.W);
      const int yy = y + (P.g.hp - P.g.h) / 2, xx = x + (P.g.wp - P.g.w) / 2;
      const int m = (yy / P.p) * P.g.nbw + xx / P.q;
      for (int c = 0; c < 3; ++c) {
        const int col = c * P.p * P.q + (yy % P.p) * P.q + xx % P.q;
        float acc = 0.0f;
        for (int r = 0; r < P.R; ++r) {
          const float uq = (float)rec[P.u_off + (size_t)r * P.g.rows + m];
          const float vq = (float)rec[P.v_off + (size_t)r * ncols + col];
          const float uf = __fadd_rn(__fmul_rn(__fsub_rn(uq, qu0), su), mu);  // dequantize, utils.py:241
          const float vf = __fadd_rn(__fmul_rn(__fsub_rn(vq, qv0), sv), mv);
          acc = __fmaf_rn(uf, vf, acc);  // u @ v.mT: ascending-r FMA chain (MKL sgemm)
        }
        o[c * hw + e] = to_u8_trunc(acc);
      }
    }
  }
}

}  // namespace lrfb

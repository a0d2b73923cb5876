// Front end: image → patch matrices, fused.
//   RGB→YCbCr      lrf/compression/utils.py:24-47   (FMA chain t0*r, fma(t1,g,.), fma(t2,b,.), + offset)
//   area 2x pool   lrf/compression/utils.py:76-95   (adaptive_avg_pool2d: row-major sum from 0, /kh, /kw)
//   reflect pad    lrf/compression/utils.py:108-132
//   patchify       lrf/compression/qmf.py:43-56     ("c (h p) (w q) -> (h w) (c p q)")
// One thread per output element of X; writes are fully coalesced, reads hit L1/L2 for the
// neighbouring patch rows.  HBM-bound: 3 B/pixel in, 6 B/pixel out (YCbCr, scale 0.5).
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {

struct PlaneGeom {
  int h, w;      // plane size before padding (chroma: after down-sampling)
  int hp, wp;    // padded size
  int top, left; // pad before
  int nbw;       // patches per row  (wp / q)
  int rows;      // M = (hp/p)*(wp/q)
};

struct FrontParams {
  int H, W;   // image size
  int p, q;   // patch size
  int ycbcr;  // 1: three single-channel planes; 0: one 3-channel RGB plane
  int n_img;
  PlaneGeom g[3];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <typename in_t>
__device__ __forceinline__ float ycc_at(const in_t* __restrict__ img, size_t hw, size_t off, int c) {
  const float t[3][3] = {{0.299f, 0.587f, 0.114f},
                         {-0.168736f, -0.331264f, 0.5f},
                         {0.5f, -0.418688f, -0.081312f}};
  const float o[3] = {0.0f, 128.0f, 128.0f};
  float r = (float)img[off], g = (float)img[hw + off], b = (float)img[2 * hw + off];
  float acc = __fmul_rn(t[c][0], r);
  acc = __fmaf_rn(t[c][1], g, acc);
  acc = __fmaf_rn(t[c][2], b, acc);
  return __fadd_rn(o[c], acc);
}

// grid.x: blocks over the elements of one plane matrix, grid.y: images (strided), plane fixed per launch
template <typename in_t>
__global__ void __launch_bounds__(256)
frontend_kernel(const in_t* __restrict__ images, float* __restrict__ xout, FrontParams P, int plane) {
  const PlaneGeom g = P.g[plane];
  const int pq = P.p * P.q;
  const int N = P.ycbcr ? pq : 3 * pq;
  const long long per_img = (long long)g.rows * N;
  const size_t hw = (size_t)P.H * P.W;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const in_t* img = images + (size_t)im * 3 * hw;
    float* x = xout + (size_t)im * per_img;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_img;
         e += (long long)gridDim.x * blockDim.x) {
      int m = (int)(e / N), n = (int)(e - (long long)m * N);
      int c = n / pq, rem = n - c * pq;
      int pi = rem / P.q, qi = rem - pi * P.q;
      int hb = m / g.nbw, wb = m - hb * g.nbw;
      int y = reflect_idx(hb * P.p + pi - g.top, g.h);
      int xx = reflect_idx(wb * P.q + qi - g.left, g.w);
      float val;
      if (!P.ycbcr) {
        val = (float)img[(size_t)c * hw + (size_t)y * P.W + xx];
      } else if (plane == 0) {
        val = ycc_at(img, hw, (size_t)y * P.W + xx, 0);
      } else {
        // adaptive_avg_pool2d window of output (y, xx): [floor(i*H/oh), ceil((i+1)*H/oh))
        int h0 = (int)(((long long)y * P.H) / g.h);
        int h1 = (int)(((long long)(y + 1) * P.H + g.h - 1) / g.h);
        int w0 = (int)(((long long)xx * P.W) / g.w);
        int w1 = (int)(((long long)(xx + 1) * P.W + g.w - 1) / g.w);
        float s = 0.0f;
        for (int yy = h0; yy < h1; ++yy)
          for (int xw = w0; xw < w1; ++xw) s = __fadd_rn(s, ycc_at(img, hw, (size_t)yy * P.W + xw, plane));
        val = __fdiv_rn(__fdiv_rn(s, (float)(h1 - h0)), (float)(w1 - w0));
      }
      x[e] = val;
    }
  }
}

}  // namespace lrfb

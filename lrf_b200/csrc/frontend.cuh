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

// ---------------------------------------------------------------------------------------------------
// Fast path: uint8 input, YCbCr, 8x8 patches, W % 16 == 0 and chroma width == W/2 (no horizontal
// padding on any plane, 2-wide pooling windows).  One thread produces 8 consecutive elements of one
// patch row: 8-byte (Y) / 16-byte (chroma) vector loads per channel, 2 x 16-byte stores.  Same
// arithmetic, operation for operation, as frontend_kernel.
// ---------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float ycc_from(float r, float g, float b, int c) {
  // c is a compile-time constant at every call site
  const float t0 = c == 0 ? 0.299f : c == 1 ? -0.168736f : 0.5f;
  const float t1 = c == 0 ? 0.587f : c == 1 ? -0.331264f : -0.418688f;
  const float t2 = c == 0 ? 0.114f : c == 1 ? 0.5f : -0.081312f;
  const float off = c == 0 ? 0.0f : 128.0f;
  float acc = __fmul_rn(t0, r);
  acc = __fmaf_rn(t1, g, acc);
  acc = __fmaf_rn(t2, b, acc);
  return __fadd_rn(off, acc);
}
// byte k of w as a float: PRMT drops the byte into the mantissa of 2^23, one exact subtraction removes the 2^23
// (two full-rate instructions instead of shift + mask + quarter-rate I2F)
__device__ __forceinline__ float byte_of(unsigned w, int k) {
#ifdef LRFB_SIM
  return (float)((w >> (8 * k)) & 0xffu);
#else
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)), 8388608.0f);
#endif
}
}  // namespace

__global__ void __launch_bounds__(256)
frontend8_luma_kernel(const unsigned char* __restrict__ images, float* __restrict__ xout, FrontParams P) {
  const PlaneGeom g = P.g[0];
  const size_t hw = (size_t)P.H * P.W;
  const int nbw = g.nbw;  // == W/8
  const long long items = (long long)g.hp * nbw;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const unsigned char* img = images + (size_t)im * 3 * hw;
    float* x = xout + (size_t)im * g.rows * 64;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (long long)gridDim.x * blockDim.x) {
      const int yy = (int)(it / nbw), wb = (int)(it - (long long)yy * nbw);
      const int y = reflect_idx(yy - g.top, g.h);
      const size_t off = (size_t)y * P.W + (size_t)wb * 8;
      const uint2 r = *reinterpret_cast<const uint2*>(img + off);
      const uint2 gg = *reinterpret_cast<const uint2*>(img + hw + off);
      const uint2 b = *reinterpret_cast<const uint2*>(img + 2 * hw + off);
      float o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[k] = ycc_from(byte_of(r.x, k), byte_of(gg.x, k), byte_of(b.x, k), 0);
        o[4 + k] = ycc_from(byte_of(r.y, k), byte_of(gg.y, k), byte_of(b.y, k), 0);
      }
      float* dst = x + ((size_t)(yy >> 3) * nbw + wb) * 64 + (yy & 7) * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// RGB planes (color_space="RGB": SVD codec, QMF on RGB), 8x8 patches, W % 8 == 0: one thread converts 8 consecutive
// pixels of one patch row of one channel (8-byte load, two 16-byte stores); X[m][c*64 + py*8 + px].
__global__ void __launch_bounds__(256)
frontend8_rgb_kernel(const unsigned char* __restrict__ images, float* __restrict__ xout, FrontParams P) {
  const PlaneGeom g = P.g[0];
  const size_t hw = (size_t)P.H * P.W;
  const int nbw = g.nbw;  // == W/8
  const long long items = 3LL * g.hp * nbw;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const unsigned char* img = images + (size_t)im * 3 * hw;
    float* x = xout + (size_t)im * g.rows * 192;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (long long)gridDim.x * blockDim.x) {
      const int wb = (int)(it % nbw);
      const long long t = it / nbw;
      const int c = (int)(t % 3), yy = (int)(t / 3);
      const int y = reflect_idx(yy - g.top, g.h);
      const uint2 v = *reinterpret_cast<const uint2*>(img + (size_t)c * hw + (size_t)y * P.W + (size_t)wb * 8);
      float* dst = x + ((size_t)(yy >> 3) * nbw + wb) * 192 + c * 64 + (yy & 7) * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(byte_of(v.x, 0), byte_of(v.x, 1), byte_of(v.x, 2), byte_of(v.x, 3));
      *reinterpret_cast<float4*>(dst + 4) = make_float4(byte_of(v.y, 0), byte_of(v.y, 1), byte_of(v.y, 2), byte_of(v.y, 3));
    }
  }
}

__global__ void __launch_bounds__(256)
frontend8_chroma_kernel(const unsigned char* __restrict__ images, float* __restrict__ xcb,
                        float* __restrict__ xcr, FrontParams P) {
  const PlaneGeom g = P.g[1];
  const size_t hw = (size_t)P.H * P.W;
  const int nbw = g.nbw;  // == W/16
  const long long items = (long long)g.hp * nbw;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const unsigned char* img = images + (size_t)im * 3 * hw;
    float* ocb = xcb + (size_t)im * g.rows * 64;
    float* ocr = xcr + (size_t)im * g.rows * 64;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (long long)gridDim.x * blockDim.x) {
      const int yy = (int)(it / nbw), wb = (int)(it - (long long)yy * nbw);
      const int cy = reflect_idx(yy - g.top, g.h);
      const int h0 = (int)(((long long)cy * P.H) / g.h);
      const int h1 = (int)(((long long)(cy + 1) * P.H + g.h - 1) / g.h);
      float sb[8], sr[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sb[j] = 0.0f, sr[j] = 0.0f;
      for (int row = h0; row < h1; ++row) {
        const size_t off = (size_t)row * P.W + (size_t)wb * 16;
        const uint4 r = *reinterpret_cast<const uint4*>(img + off);
        const uint4 gg = *reinterpret_cast<const uint4*>(img + hw + off);
        const uint4 b = *reinterpret_cast<const uint4*>(img + 2 * hw + off);
        const unsigned rw[4] = {r.x, r.y, r.z, r.w}, gw[4] = {gg.x, gg.y, gg.z, gg.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int px = 2 * j + s;
            const float fr = byte_of(rw[px >> 2], px & 3), fg = byte_of(gw[px >> 2], px & 3),
                        fb = byte_of(bw[px >> 2], px & 3);
            sb[j] = __fadd_rn(sb[j], ycc_from(fr, fg, fb, 1));
            sr[j] = __fadd_rn(sr[j], ycc_from(fr, fg, fb, 2));
          }
        }
      }
      const float kh = (float)(h1 - h0);
      if (h1 - h0 == 2) {  // the usual window: dividing by 2 is an exact scaling, two multiplies give the same bits
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sb[j] = __fmul_rn(__fmul_rn(sb[j], 0.5f), 0.5f);
          sr[j] = __fmul_rn(__fmul_rn(sr[j], 0.5f), 0.5f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sb[j] = __fmul_rn(__fdiv_rn(sb[j], kh), 0.5f);
          sr[j] = __fmul_rn(__fdiv_rn(sr[j], kh), 0.5f);
        }
      }
      const size_t o = ((size_t)(yy >> 3) * nbw + wb) * 64 + (yy & 7) * 8;
      *reinterpret_cast<float4*>(ocb + o) = make_float4(sb[0], sb[1], sb[2], sb[3]);
      *reinterpret_cast<float4*>(ocb + o + 4) = make_float4(sb[4], sb[5], sb[6], sb[7]);
      *reinterpret_cast<float4*>(ocr + o) = make_float4(sr[0], sr[1], sr[2], sr[3]);
      *reinterpret_cast<float4*>(ocr + o + 4) = make_float4(sr[4], sr[5], sr[6], sr[7]);
    }
  }
}

// Fused variant for H % 16 == 0 and W % 16 == 0 (no padding on any plane, 2 x 2 pooling windows): the image is
// read ONCE for all three planes, and the patch rows are staged in shared memory so that global memory sees only
// fully coalesced accesses.  A CTA owns a 16-row x 256-pixel tile: thread (cy, wp) reads the 2-row x 8-pixel strip
// of the three channels (3 x 2 x 8-byte loads; a warp covers 256 contiguous bytes of an image row), produces its
// 2 x 8 luma values and the 4 Cb / 4 Cr values pooled from the strip, and scatters them into the tile's patch
// layout (patch stride 68 floats: conflict-free 16-byte stores).  The tile's 2 x 32 luma patches and 16 + 16 chroma
// patches are each contiguous in X, so the write-out is plain consecutive float4 stores.  Same arithmetic,
// operation for operation, as the kernels above.
constexpr int kTilePatchStride = 68;
__global__ void __launch_bounds__(256)
frontend8_fused_kernel(const unsigned char* __restrict__ images, float* __restrict__ xy, float* __restrict__ xcb,
                       float* __restrict__ xcr, FrontParams P) {
  __align__(16) __shared__ float s_lum[2 * 32 * kTilePatchStride];
  __align__(16) __shared__ float s_cb[16 * kTilePatchStride];
  __align__(16) __shared__ float s_cr[16 * kTilePatchStride];
  const PlaneGeom gl = P.g[0], gc = P.g[1];
  const size_t hw = (size_t)P.H * P.W;
  const int nbl = gl.nbw, nbc = gc.nbw;  // W/8 and W/16 patches per patch row
  const int tid = threadIdx.x, wp = tid & 31, cy = tid >> 5;
  const int y0 = blockIdx.y * 16, wp0 = blockIdx.x * 32;  // first image row / first luma patch column of the tile
  const int n_lp = min(32, nbl - wp0), n_cp = n_lp >> 1;  // patches of this tile that exist
  for (int im = blockIdx.z; im < P.n_img; im += gridDim.z) {
    const unsigned char* img = images + (size_t)im * 3 * hw;
    __syncthreads();  // previous image's write-out is done with the staging buffers
    if (wp < n_lp) {
      float sb[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int ly = 2 * cy + dy;
        const size_t off = (size_t)(y0 + ly) * P.W + (size_t)(wp0 + wp) * 8;
        const uint2 r = *reinterpret_cast<const uint2*>(img + off);
        const uint2 gg = *reinterpret_cast<const uint2*>(img + hw + off);
        const uint2 b = *reinterpret_cast<const uint2*>(img + 2 * hw + off);
        const unsigned rw[2] = {r.x, r.y}, gw[2] = {gg.x, gg.y}, bw[2] = {b.x, b.y};
        float lum[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int px = 2 * j + s;
            const float fr = byte_of(rw[px >> 2], px & 3), fg = byte_of(gw[px >> 2], px & 3),
                        fb = byte_of(bw[px >> 2], px & 3);
            lum[px] = ycc_from(fr, fg, fb, 0);
            sb[j] = __fadd_rn(sb[j], ycc_from(fr, fg, fb, 1));
            sr[j] = __fadd_rn(sr[j], ycc_from(fr, fg, fb, 2));
          }
        }
        float* d = &s_lum[((ly >> 3) * 32 + wp) * kTilePatchStride + (ly & 7) * 8];
        *reinterpret_cast<float4*>(d) = make_float4(lum[0], lum[1], lum[2], lum[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(lum[4], lum[5], lum[6], lum[7]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // (s / 2) / 2: scaling by a power of two is exact, so two multiplies give the same bits
        sb[j] = __fmul_rn(__fmul_rn(sb[j], 0.5f), 0.5f);
        sr[j] = __fmul_rn(__fmul_rn(sr[j], 0.5f), 0.5f);
      }
      const int co = (wp >> 1) * kTilePatchStride + cy * 8 + (wp & 1) * 4;
      *reinterpret_cast<float4*>(&s_cb[co]) = make_float4(sb[0], sb[1], sb[2], sb[3]);
      *reinterpret_cast<float4*>(&s_cr[co]) = make_float4(sr[0], sr[1], sr[2], sr[3]);
    }
    __syncthreads();
    // write-out: 2 runs of n_lp luma patches, 1 run of n_cp patches per chroma plane, all contiguous in X
    float* oy = xy + (size_t)im * gl.rows * 64;
    float* ocb = xcb + (size_t)im * gc.rows * 64;
    float* ocr = xcr + (size_t)im * gc.rows * 64;
    if (n_lp == 32) {  // full tile (every tile when W % 256 == 0): index math in shifts, 4 + 2 stores per thread
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = tid + t * 256, pr = i >> 9, patch = (i >> 4) & 31, q = i & 15;
        const float4 v = *reinterpret_cast<const float4*>(&s_lum[(pr * 32 + patch) * kTilePatchStride + q * 4]);
        *reinterpret_cast<float4*>(oy + ((size_t)((y0 >> 3) + pr) * nbl + wp0 + patch) * 64 + q * 4) = v;
      }
      const int patch = tid >> 4, q = tid & 15;
      const size_t co = ((size_t)(y0 >> 4) * nbc + (wp0 >> 1) + patch) * 64 + q * 4;
      *reinterpret_cast<float4*>(ocb + co) = *reinterpret_cast<const float4*>(&s_cb[patch * kTilePatchStride + q * 4]);
      *reinterpret_cast<float4*>(ocr + co) = *reinterpret_cast<const float4*>(&s_cr[patch * kTilePatchStride + q * 4]);
    } else {
      for (int i = tid; i < 2 * n_lp * 16; i += 256) {
        const int pr = i / (n_lp * 16), rem = i - pr * n_lp * 16, patch = rem >> 4, q = rem & 15;
        const float4 v = *reinterpret_cast<const float4*>(&s_lum[(pr * 32 + patch) * kTilePatchStride + q * 4]);
        *reinterpret_cast<float4*>(oy + ((size_t)((y0 >> 3) + pr) * nbl + wp0 + patch) * 64 + q * 4) = v;
      }
      for (int i = tid; i < 2 * n_cp * 16; i += 256) {
        const int pl = i / (n_cp * 16), rem = i - pl * n_cp * 16, patch = rem >> 4, q = rem & 15;
        const float4 v = *reinterpret_cast<const float4*>(&(pl ? s_cr : s_cb)[patch * kTilePatchStride + q * 4]);
        *reinterpret_cast<float4*>((pl ? ocr : ocb) + ((size_t)(y0 >> 4) * nbc + (wp0 >> 1) + patch) * 64 + q * 4) = v;
      }
    }
  }
}

}  // namespace lrfb

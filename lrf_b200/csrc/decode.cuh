// Fused decoder: int8 factors → uint8 RGB image, one thread per output pixel.
//   QMF.reconstruct (u @ v.mT)   lrf/factorization/qmf.py:216-223   exact small-integer arithmetic
//   depatchify                   lrf/compression/qmf.py:59-75
//   unpad_image                  lrf/compression/utils.py:135-153   (start = (Hp-H)//2)
//   chroma_upsampling (nearest)  lrf/compression/utils.py:98-105    src = min(floorf(dst*in/out), in-1)
//   ycbcr_to_rgb                 lrf/compression/utils.py:50-73     FMA chain on (ycc + offset)
//   to_dtype(uint8)              lrf/compression/utils.py:156-182   clamp then truncate
// HBM-bound: ~0.08 B/pixel of factors in, 3 B/pixel out.  Also the exact per-image SSE used by PSNR
// (lrf/utils/metrics.py:24-35, :57-71).
#pragma once
#include "frontend.cuh"

namespace lrfb {

struct DecodeParams {
  int H, W, p, q, ycbcr, n_img;
  PlaneGeom g[3];
  int rank[3];
  long long record_bytes;
  long long u_off[3], v_off[3];
};

// clamp to [0, 255], then truncate.  On the device the truncation is a round-toward-zero add onto 2^23 (the integer
// part lands in the low mantissa bits): full-rate FADD instead of the quarter-rate F2I that bounded the decoder.
__device__ __forceinline__ unsigned to_u8_bits(float v) {  // result in the low byte
  v = fminf(fmaxf(v, 0.0f), 255.0f);
#ifdef LRFB_SIM
  return (unsigned)(int)v;
#else
  return __float_as_uint(__fadd_rz(v, 8388608.0f)) & 0xffu;
#endif
}
__device__ __forceinline__ unsigned char to_u8_trunc(float v) { return (unsigned char)to_u8_bits(v); }

__device__ __forceinline__ float plane_value(const int8_t* __restrict__ rec, const DecodeParams& P, int c,
                                             int chan, int y, int x) {
  const PlaneGeom& g = P.g[c];
  const int yy = y + (g.hp - g.h) / 2, xx = x + (g.wp - g.w) / 2;
  const int m = (yy / P.p) * g.nbw + xx / P.q;
  const int col = chan * P.p * P.q + (yy % P.p) * P.q + xx % P.q;
  const int ncols = P.ycbcr ? P.p * P.q : 3 * P.p * P.q;
  const int8_t* u = rec + P.u_off[c];
  const int8_t* v = rec + P.v_off[c];
  float acc = 0.0f;
  for (int r = 0; r < P.rank[c]; ++r)
    acc = __fadd_rn(acc, __fmul_rn((float)u[(size_t)r * g.rows + m], (float)v[(size_t)r * ncols + col]));
  return acc;
}

__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  if (out == in) return dst;
  if (out == 2 * in) return dst >> 1;
  float scale = __fdiv_rn((float)in, (float)out);
  int s = (int)floorf(__fmul_rn((float)dst, scale));
  return min(s, in - 1);
}

__global__ void __launch_bounds__(256)
qmf_decode_kernel(const int8_t* __restrict__ factors, unsigned char* __restrict__ out, DecodeParams P) {
  const float t[3][3] = {{1.0f, 0.0f, 1.40200f}, {1.0f, -0.344136f, -0.714136f}, {1.0f, 1.77200f, 0.0f}};
  const size_t hw = (size_t)P.H * P.W;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const int8_t* rec = factors + (size_t)im * P.record_bytes;
    unsigned char* o = out + (size_t)im * 3 * hw;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < hw; e += (size_t)gridDim.x * blockDim.x) {
      const int y = (int)(e / P.W), x = (int)(e - (size_t)y * P.W);
      if (!P.ycbcr) {
        for (int c = 0; c < 3; ++c) o[c * hw + e] = to_u8_trunc(plane_value(rec, P, 0, c, y, x));
        continue;
      }
      const int sy = nearest_src(y, P.g[1].h, P.H), sx = nearest_src(x, P.g[1].w, P.W);
      float ycc[3];
      ycc[0] = __fadd_rn(plane_value(rec, P, 0, 0, y, x), 0.0f);
      ycc[1] = __fadd_rn(plane_value(rec, P, 1, 0, sy, sx), -128.0f);
      ycc[2] = __fadd_rn(plane_value(rec, P, 2, 0, sy, sx), -128.0f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = __fmul_rn(t[c][0], ycc[0]);
        acc = __fmaf_rn(t[c][1], ycc[1], acc);
        acc = __fmaf_rn(t[c][2], ycc[2], acc);
        o[c * hw + e] = to_u8_trunc(acc);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// patch=False streams (lrf/compression/qmf.py:303-311 RGB, :339-344 YCbCr): every plane is U V^T of whole-channel
// factors kept row-major with their leading batch dimension, u [n_img][(3)][h][R], v [n_img][(3)][w][R].
// ---------------------------------------------------------------------------------------------------
struct PlanesParams {
  int H, W, ch, cw, ycbcr, n_img;
  int rank[3];
  const int8_t* u[3];
  const int8_t* v[3];
};

__device__ __forceinline__ float uv_dot(const int8_t* __restrict__ u, const int8_t* __restrict__ v, int R) {
  float acc = 0.0f;  // small integers: exact in f32 in any order
  for (int r = 0; r < R; ++r) acc = __fadd_rn(acc, __fmul_rn((float)u[r], (float)v[r]));
  return acc;
}

__global__ void __launch_bounds__(256)
qmf_decode_planes_kernel(unsigned char* __restrict__ out, PlanesParams P) {
  const float t[3][3] = {{1.0f, 0.0f, 1.40200f}, {1.0f, -0.344136f, -0.714136f}, {1.0f, 1.77200f, 0.0f}};
  const size_t hw = (size_t)P.H * P.W;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    unsigned char* o = out + (size_t)im * 3 * hw;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < hw; e += (size_t)gridDim.x * blockDim.x) {
      const int y = (int)(e / P.W), x = (int)(e - (size_t)y * P.W);
      if (!P.ycbcr) {
        const int R = P.rank[0];
        for (int c = 0; c < 3; ++c)
          o[c * hw + e] = to_u8_trunc(uv_dot(P.u[0] + (((size_t)im * 3 + c) * P.H + y) * R,
                                            P.v[0] + (((size_t)im * 3 + c) * P.W + x) * R, R));
        continue;
      }
      const int sy = nearest_src(y, P.ch, P.H), sx = nearest_src(x, P.cw, P.W);
      float ycc[3];
      ycc[0] = __fadd_rn(uv_dot(P.u[0] + ((size_t)im * P.H + y) * P.rank[0], P.v[0] + ((size_t)im * P.W + x) * P.rank[0], P.rank[0]), 0.0f);
      ycc[1] = __fadd_rn(uv_dot(P.u[1] + ((size_t)im * P.ch + sy) * P.rank[1], P.v[1] + ((size_t)im * P.cw + sx) * P.rank[1], P.rank[1]), -128.0f);
      ycc[2] = __fadd_rn(uv_dot(P.u[2] + ((size_t)im * P.ch + sy) * P.rank[2], P.v[2] + ((size_t)im * P.cw + sx) * P.rank[2], P.rank[2]), -128.0f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = __fmul_rn(t[c][0], ycc[0]);
        acc = __fmaf_rn(t[c][1], ycc[1], acc);
        acc = __fmaf_rn(t[c][2], ycc[2], acc);
        o[c * hw + e] = to_u8_trunc(acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast path: YCbCr, 8x8 patches, W % 16 == 0, chroma width == W/2.  One thread reconstructs 8
// consecutive pixels of one image row: they lie in one luma patch row and 4 consecutive elements of one
// chroma patch row, so the factors are read as 8-byte / 4-byte vectors and each colour plane gets one
// 8-byte store.  Same arithmetic as qmf_decode_kernel.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sbyte_of(unsigned w, int k) { return (float)(signed char)((w >> (8 * k)) & 0xffu); }

__global__ void __launch_bounds__(256)
qmf_decode8_kernel(const int8_t* __restrict__ factors, unsigned char* __restrict__ out, DecodeParams P) {
  const float t[3][3] = {{1.0f, 0.0f, 1.40200f}, {1.0f, -0.344136f, -0.714136f}, {1.0f, 1.77200f, 0.0f}};
  const size_t hw = (size_t)P.H * P.W;
  const int segs = P.W / 8;
  const long long items = (long long)P.H * segs;
  const PlaneGeom gy = P.g[0], gc = P.g[1];
  const int shy = (gy.hp - gy.h) / 2, shc = (gc.hp - gc.h) / 2;
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const int8_t* rec = factors + (size_t)im * P.record_bytes;
    unsigned char* o = out + (size_t)im * 3 * hw;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (long long)gridDim.x * blockDim.x) {
      const int y = (int)(it / segs), sg = (int)(it - (long long)y * segs);
      float py[8], pcb[4], pcr[4];
      {  // luma
        const int yy = y + shy, m = (yy >> 3) * gy.nbw + sg, col = (yy & 7) * 8;
        const int8_t* u = rec + P.u_off[0];
        const int8_t* v = rec + P.v_off[0];
#pragma unroll
        for (int j = 0; j < 8; ++j) py[j] = 0.0f;
        for (int r = 0; r < P.rank[0]; ++r) {
          const float ur = (float)u[(size_t)r * gy.rows + m];
          const uint2 vv = *reinterpret_cast<const uint2*>(v + (size_t)r * 64 + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            py[j] = __fadd_rn(py[j], __fmul_rn(ur, sbyte_of(vv.x, j)));
            py[4 + j] = __fadd_rn(py[4 + j], __fmul_rn(ur, sbyte_of(vv.y, j)));
          }
        }
      }
      {  // chroma: pixels 8*sg .. 8*sg+7 use chroma columns 4*sg .. 4*sg+3
        const int sy = nearest_src(y, gc.h, P.H);
        const int yy = sy + shc, cx0 = 4 * sg;
        const int m = (yy >> 3) * gc.nbw + (cx0 >> 3), col = (yy & 7) * 8 + (cx0 & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j) pcb[j] = 0.0f, pcr[j] = 0.0f;
        for (int r = 0; r < P.rank[1]; ++r) {
          const float ur = (float)rec[P.u_off[1] + (size_t)r * gc.rows + m];
          const unsigned vv = *reinterpret_cast<const unsigned*>(rec + P.v_off[1] + (size_t)r * 64 + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) pcb[j] = __fadd_rn(pcb[j], __fmul_rn(ur, sbyte_of(vv, j)));
        }
        for (int r = 0; r < P.rank[2]; ++r) {
          const float ur = (float)rec[P.u_off[2] + (size_t)r * gc.rows + m];
          const unsigned vv = *reinterpret_cast<const unsigned*>(rec + P.v_off[2] + (size_t)r * 64 + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) pcr[j] = __fadd_rn(pcr[j], __fmul_rn(ur, sbyte_of(vv, j)));
        }
      }
      unsigned lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y0 = __fadd_rn(py[j], 0.0f);
        const float cb = __fadd_rn(pcb[j >> 1], -128.0f), cr = __fadd_rn(pcr[j >> 1], -128.0f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float acc = __fmul_rn(t[c][0], y0);
          acc = __fmaf_rn(t[c][1], cb, acc);
          acc = __fmaf_rn(t[c][2], cr, acc);
          const unsigned b = to_u8_trunc(acc);
          if (j < 4) lo[c] |= b << (8 * j);
          else hi[c] |= b << (8 * (j - 4));
        }
      }
      const size_t off = (size_t)y * P.W + (size_t)sg * 8;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint2 w;
        w.x = lo[c], w.y = hi[c];
        *reinterpret_cast<uint2*>(o + c * hw + off) = w;
      }
    }
  }
}

#ifndef LRFB_SIM
// ---------------------------------------------------------------------------------------------------
// Faster path for the unpadded geometry (H, W multiples of 16, chroma exactly half size) and ranks <= 4: one thread
// reconstructs an 8-pixel x 2-row strip.  u @ v.T on small integers is exact in any arithmetic (|sum| <= 4*128*128),
// so it is done as int8 dot products: the 4 ranks of U in one word, the V bytes of the 4 ranks transposed into one
// word per pixel (PRMT), one DP4A per pixel instead of 4 x (convert, multiply, add).  The two rows share the luma U
// word and the whole chroma reconstruction (nearest up-sampling).  Colour transform and u8 conversion as above.
// ---------------------------------------------------------------------------------------------------
// clamp to [0, 255] and truncate in one instruction (PTX: float-to-integer conversions saturate to the destination range)
__device__ __forceinline__ unsigned f32_to_u8_sat(float x) {
  unsigned r;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void transpose4x4_bytes(unsigned a, unsigned b, unsigned c, unsigned d, unsigned (&w)[4]) {
  const unsigned t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(a, b, 0x7362);
  const unsigned u0 = __byte_perm(c, d, 0x5140), u1 = __byte_perm(c, d, 0x7362);
  w[0] = __byte_perm(t0, u0, 0x5410), w[1] = __byte_perm(t0, u0, 0x7632);
  w[2] = __byte_perm(t1, u1, 0x5410), w[3] = __byte_perm(t1, u1, 0x7632);
}

__global__ void __launch_bounds__(256)
qmf_decode8x2_kernel(const int8_t* __restrict__ factors, unsigned char* __restrict__ out, DecodeParams P) {
  const float t[3][3] = {{1.0f, 0.0f, 1.40200f}, {1.0f, -0.344136f, -0.714136f}, {1.0f, 1.77200f, 0.0f}};
  // V of the three planes, transposed once per image into DP4A operands: word [plane][position in the 8 x 8 patch]
  // holds the (up to) four rank bytes of that position; every thread of the block reads the same words (broadcast)
  __shared__ __align__(16) unsigned vt[3][64];
  const size_t hw = (size_t)P.H * P.W;
  const int segs = P.W / 8;
  const long long items = (long long)(P.H / 8) * segs;  // one item = one luma patch (8 rows x 8 pixels)
  const PlaneGeom gy = P.g[0], gc = P.g[1];
  for (int im = blockIdx.y; im < P.n_img; im += gridDim.y) {
    const int8_t* rec = factors + (size_t)im * P.record_bytes;
    unsigned char* o = out + (size_t)im * 3 * hw;
    __syncthreads();
    if (threadIdx.x < 192) {
      const int pl = threadIdx.x >> 6, pos = threadIdx.x & 63;
      unsigned w = 0;
      for (int r = 0; r < P.rank[pl]; ++r) w |= (unsigned)(unsigned char)rec[P.v_off[pl] + (size_t)r * 64 + pos] << (8 * r);
      vt[pl][pos] = w;
    }
    __syncthreads();
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (long long)gridDim.x * blockDim.x) {
      const int pr = (int)(it / segs), sg = (int)(it - (long long)pr * segs);
      // The 8 rows of the patch share the luma U word; their 4 chroma rows lie in one chroma patch and share its U words:
      // the byte gathers from the fiber-major factors and the index arithmetic are paid once per 64 pixels.
      const int m = pr * gy.nbw + sg;
      const int8_t* u = rec + P.u_off[0];
      unsigned uw = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < P.rank[0]) uw |= (unsigned)(unsigned char)u[(size_t)r * gy.rows + m] << (8 * r);
      const int cx0 = 4 * sg, mc = (pr >> 1) * gc.nbw + (cx0 >> 3);
      unsigned uwc[2] = {0, 0};
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < P.rank[1 + pl])
            uwc[pl] |= (unsigned)(unsigned char)rec[P.u_off[1 + pl] + (size_t)r * gc.rows + mc] << (8 * r);
      unsigned char* orow = o + (size_t)(8 * pr) * P.W + (size_t)sg * 8;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int cy = 4 * pr + q;
        // ---- chroma: 4 columns 4*sg .. 4*sg+3 of chroma row cy, both planes ----
        const int cpos = (cy & 7) * 8 + (cx0 & 7);
        const uint4 wb = *reinterpret_cast<const uint4*>(&vt[1][cpos]), wr = *reinterpret_cast<const uint4*>(&vt[2][cpos]);
        const unsigned wcb[4] = {wb.x, wb.y, wb.z, wb.w}, wcr[4] = {wr.x, wr.y, wr.z, wr.w};
        float cb[4], cr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cb[j] = __fadd_rn((float)__dp4a((int)wcb[j], (int)uwc[0], 0), -128.0f);
          cr[j] = __fadd_rn((float)__dp4a((int)wcr[j], (int)uwc[1], 0), -128.0f);
        }
        // ---- luma rows 2q and 2q+1 of the patch ----
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int row = 2 * q + dy;
          const uint4 w0 = *reinterpret_cast<const uint4*>(&vt[0][row * 8]), w1 = *reinterpret_cast<const uint4*>(&vt[0][row * 8 + 4]);
          const unsigned wy[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          unsigned px[3][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float yv = (float)__dp4a((int)wy[j], (int)uw, 0);
            // t[c][0] = 1 and the zero coefficients drop out exactly (1*y = y; fma(0, c, a) = a up to the sign of a
            // zero, which the conversion maps to 0 either way); the saturating conversion is clamp + truncate
            px[0][j] = f32_to_u8_sat(__fmaf_rn(t[0][2], cr[j >> 1], yv));
            px[1][j] = f32_to_u8_sat(__fmaf_rn(t[1][2], cr[j >> 1], __fmaf_rn(t[1][1], cb[j >> 1], yv)));
            px[2][j] = f32_to_u8_sat(__fmaf_rn(t[2][1], cb[j >> 1], yv));
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {  // low bytes of 4 words -> one word: 3 PRMT
            const unsigned lo = __byte_perm(__byte_perm(px[c][0], px[c][1], 0x0040), __byte_perm(px[c][2], px[c][3], 0x0040), 0x5410);
            const unsigned hi = __byte_perm(__byte_perm(px[c][4], px[c][5], 0x0040), __byte_perm(px[c][6], px[c][7], 0x0040), 0x5410);
            *reinterpret_cast<uint2*>(orow + c * hw + (size_t)row * P.W) = make_uint2(lo, hi);
          }
        }
      }
    }
  }
}
#endif  // LRFB_SIM

// exact sum of squared differences per image: integer arithmetic, order-independent
__global__ void __launch_bounds__(256)
sse_u8_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, long long per_img,
              int n_img, unsigned long long* __restrict__ sse) {
  __shared__ unsigned long long warp_sums[8];
  for (int im = blockIdx.y; im < n_img; im += gridDim.y) {
    const unsigned char* pa = a + (size_t)im * per_img;
    const unsigned char* pb = b + (size_t)im * per_img;
    unsigned long long s = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_img;
         e += (long long)gridDim.x * blockDim.x) {
      int d = (int)pa[e] - (int)pb[e];
      s += (unsigned long long)(d * d);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_sums[w];
      atomicAdd(&sse[im], tot);
    }
  }
}

}  // namespace lrfb

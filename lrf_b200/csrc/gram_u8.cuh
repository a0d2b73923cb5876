// Exact Gram matrix G = X^T X for N = 192 patch matrices whose entries are integers in [0, 255] — the RGB planes of
// uint8 images (color_space="RGB": the SVD baseline codec, QMF on RGB).  The bytes themselves are the operands of
// tcgen05.mma kind::i8 (u8 x u8 -> s32, exact for <= 66 000 rows): one slice instead of the four of gram64_i8_kernel.
//   D1[128 x 192] += X[:, 0..128)^T  X,   D2[128 x 192] += X[:, 64..192)^T  X      per 32 rows
// (two M = 128 products with overlapping row ranges rather than an M = 64 one: plain lane = row accumulator layout).
// Both operands MN-major, no swizzle (core matrix = 8 rows x 16 columns).  Warps 0-7 convert and stage 64-row tiles
// (4 stages, loads of the next tile in flight), warp 8 issues the MMAs, warps 0-3 write G from tensor memory.
#pragma once
#include "gram_i8.cuh"

#ifndef LRFB_SIM

namespace lrfb {

constexpr int kU8N = 192;
constexpr int kU8TileRows = 64;
constexpr int kU8Sbo = (kU8TileRows / 8) * 128 + 32;   // MN-core stride: 8 K-cores of 128 B + bank padding
constexpr int kU8StageBytes = (kU8N / 16) * kU8Sbo;    // 12 MN-cores: 12 672 B
constexpr int kU8Stages = 4;

struct GramU8Smem {
  unsigned char stage[kU8Stages][kU8StageBytes];
  unsigned long long full[kU8Stages], empty[kU8Stages], done;
  unsigned tmem_base;
};

// grid = (row splits, matrices).  Gout[(mat*n_split + split)][192][192] f64.
__global__ void __launch_bounds__(kI8Threads, 1)
gram192_u8_kernel(const float* __restrict__ X, long long x_stride, int M, double* __restrict__ Gout, int n_split) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  GramU8Smem& sm = *reinterpret_cast<GramU8Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mat = blockIdx.y, split = blockIdx.x;
  const float* x = X + (size_t)mat * x_stride;
  const int n_tiles = (M + kU8TileRows - 1) / kU8TileRows;
  const int my_tiles = n_tiles > split ? (n_tiles - split + n_split - 1) / n_split : 0;

  if (tid == 0) {
    for (int s = 0; s < kU8Stages; ++s) mbar_init(&sm.full[s], kI8ProdWarps * 32), mbar_init(&sm.empty[s], 1);
    mbar_init(&sm.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kI8ProdWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = sm.tmem_base;

  if (warp < kI8ProdWarps) {
    // ---------------- producers: integer-valued f32 rows -> bytes in the core-matrix layout ----------------
    constexpr int IPT = kU8TileRows * (kU8N / 4) / (kI8ProdWarps * 32);  // float4 items per thread per tile (12)
    float4 ring[2][IPT];
    auto fetch = [&](float4 (&dst)[IPT], int it) {
      const int r0 = (split + it * n_split) * kU8TileRows;
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        const int e = tid + q * (kI8ProdWarps * 32);
        const int row = e / (kU8N / 4), c4 = (e - row * (kU8N / 4)) * 4;
        dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < my_tiles && r0 + row < M) dst[q] = *reinterpret_cast<const float4*>(x + (size_t)(r0 + row) * kU8N + c4);
      }
    };
    auto convert = [&](const float4 (&cur)[IPT], int it) {
      const int s = it % kU8Stages;
      if (it >= kU8Stages) mbar_wait(&sm.empty[s], ((it / kU8Stages) - 1) & 1);
      unsigned char* st = sm.stage[s];
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        const int e = tid + q * (kI8ProdWarps * 32);
        const int row = e / (kU8N / 4), c4 = (e - row * (kU8N / 4)) * 4;
        const float4 v = cur[q];
        const unsigned b0 = __float2uint_rn(v.x), b1 = __float2uint_rn(v.y), b2 = __float2uint_rn(v.z), b3 = __float2uint_rn(v.w);
        const unsigned w = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
        const unsigned off = (c4 >> 4) * kU8Sbo + (row >> 3) * 128 + (row & 7) * 16 + (c4 & 15);
        *reinterpret_cast<unsigned*>(st + off) = w;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&sm.full[s]);
    };
    fetch(ring[0], 0);
    for (int it = 0; it < my_tiles; it += 2) {
      fetch(ring[1], it + 1);
      convert(ring[0], it);
      if (it + 1 < my_tiles) {
        fetch(ring[0], it + 2);
        convert(ring[1], it + 1);
      }
    }
  } else if (lane == 0) {
    // ---------------- MMA issuer: D = s32, A = B = u8, both MN-major, N = 192, M = 128 ----------------
    const unsigned idesc = (2u << 4) | (1u << 15) | (1u << 16) | ((192u >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % kU8Stages;
      mbar_wait(&sm.full[s], (it / kU8Stages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned base = smem_u32(sm.stage[s]);
#pragma unroll
      for (int j = 0; j < kU8TileRows / 32; ++j) {
        const unsigned long long bdesc = umma_desc(base + j * 512, 128, kU8Sbo);                  // columns 0..191
        const unsigned long long a2 = umma_desc(base + 4 * kU8Sbo + j * 512, 128, kU8Sbo);       // columns 64..191
        const unsigned acc = (it > 0 || j > 0) ? 1u : 0u;
        umma_i8(tmem + 0, bdesc, bdesc, idesc, acc);
        umma_i8(tmem + 192, a2, bdesc, idesc, acc);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.done)) : "memory");
  }

  // ---------------- epilogue: rows 0..127 from D1 (lane = row), rows 128..191 from lanes 64..127 of D2 ----------------
  double* g = Gout + ((size_t)mat * n_split + split) * (size_t)(kU8N * kU8N);
  if (warp < 4) {
    const int r = warp * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(&sm.done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned lane_addr = tmem + ((unsigned)(warp * 32) << 16);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int row = half == 0 ? r : 64 + r;
        const bool write = half == 0 || r >= 64;
#pragma unroll 1
        for (int c0 = 0; c0 < kU8N; c0 += 16) {
          unsigned v[16];
          tmem_ld16(lane_addr + half * 192 + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (write) {
#pragma unroll
            for (int j = 0; j < 16; ++j) g[(size_t)row * kU8N + c0 + j] = (double)v[j];
          }
        }
      }
    } else {
      for (int j = 0; j < kU8N; ++j) {
        g[(size_t)r * kU8N + j] = 0.0;
        if (r >= 64) g[(size_t)(64 + r) * kU8N + j] = 0.0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kI8ProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace lrfb

#endif  // LRFB_SIM

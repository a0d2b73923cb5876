// Exact Gram matrix G = X^T X on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// TMEM) for N = 64 patch matrices whose entries lie in [0, 256) — every plane the uint8 front end produces.
//
// x * 2^24 truncated to an unsigned 32-bit word is an exact Q8.24 image of the f32 value for x >= 0.5 (below
// that the dropped bits are < 2^-24 absolute); its four BYTES are four uint8 slices S_0..S_3 with
// x = sum_a S_a 2^(-8a).  Then G = sum_{a,b} 2^(-8(a+b)) S_a^T S_b, and each S_a^T S_b is an integer matrix
// product that the int8 tensor cores compute exactly (int32 accumulation, < 2^32 for <= 66 000 rows).
// Two MMAs per 32 rows cover all 16 slice pairs at once:
//     D1[128 x 256] += [S_0 ; S_1]^T-stack (M = 128)  x  [S_0 | S_1 | S_2 | S_3] (N = 256),   K = 32 rows
//     D2[128 x 256] += [S_2 ; S_3]^T-stack             x  the same B
// (both operands "MN-major", no swizzle: a core matrix is 8 rows x 16 columns = 128 contiguous bytes).
// The epilogue reads the 128 x 512 int32 accumulators with tcgen05.ld and combines them in f64 — the only
// rounding in the whole Gram is that final 16-term sum (1e-16 relative), so the SVD initialisation keeps
// the FP64-grade accuracy it needs at a fraction of the FP64 cost.
//
// Roles: warps 0-7 convert/stage tiles of 128 rows (2 stages; the global loads of the next tile are in flight
// while the current one is converted), warps 0-3 also run the epilogue, one lane of warp 8 issues the MMAs; mbarriers full[2]/empty[2]/done; TMEM: all 512 columns.
#pragma once
#include "lrfb_common.cuh"

#ifndef LRFB_SIM

namespace lrfb {

constexpr int kI8TileRows = 128;                  // rows per stage (4 MMA k-steps of 32)
constexpr int kI8Sbo = 2048 + 32;                 // MN-core stride: 16 K-cores of 128 B + 32 B so the staging stores hit 32 banks
constexpr int kI8SliceBytes = 4 * kI8Sbo;         // one uint8 slice of a stage (4 MN-cores of 16 columns)
constexpr int kI8StageBytes = 4 * kI8SliceBytes;  // 33 280 B
constexpr int kI8ProdWarps = 8;                   // converting/staging warps (warps 0-3 also run the epilogue)
constexpr int kI8Threads = (kI8ProdWarps + 1) * 32;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// shared-memory matrix descriptor: MN-major, no swizzle, LBO = K-core stride, SBO = MN-core stride
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base offset 0, layout type 0 (no swizzle)
}

__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc,
                                        unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

#ifdef LRFB_I8_DEBUG
__device__ int g_i8_mode = 0;  // probe only: 1 = skip conversion work, 2 = skip MMAs
#define I8_MODE g_i8_mode
#else
#define I8_MODE 0
#endif

struct GramI8Smem {
  unsigned char stage[2][kI8StageBytes];  // reused as 128 x 64 f64 partials in the epilogue
  unsigned long long full[2], empty[2], done;
  unsigned tmem_base;
};

// grid = (row splits, matrices).  Gout[(mat*n_split + split)][64][64] f64.
__global__ void __launch_bounds__(kI8Threads, 1)
gram64_i8_kernel(const float* __restrict__ X, long long x_stride, int M, double* __restrict__ Gout, int n_split) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  GramI8Smem& sm = *reinterpret_cast<GramI8Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mat = blockIdx.y, split = blockIdx.x;
  const float* x = X + (size_t)mat * x_stride;
  const int n_tiles = (M + kI8TileRows - 1) / kI8TileRows;
  const int my_tiles = n_tiles > split ? (n_tiles - split + n_split - 1) / n_split : 0;

  if (tid == 0) {
    mbar_init(&sm.full[0], kI8ProdWarps * 32), mbar_init(&sm.full[1], kI8ProdWarps * 32);
    mbar_init(&sm.empty[0], 1), mbar_init(&sm.empty[1], 1);
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
    // ---------------- producers: f32 rows -> Q8.24 -> 4 byte planes in the canonical core-matrix layout ----------
    constexpr int IPT = kI8TileRows * 16 / (kI8ProdWarps * 32);  // float4 items per thread per tile (8)
    // Register ring of three tiles: the loads of tiles it+1 and it+2 are in flight while tile it is converted (one tile
    // of look-ahead left the kernel at 49 % of DRAM throughput with long-scoreboard stalls on top: a tile is converted
    // faster than a load round trip under load).
    float4 ring[3][IPT];
    auto fetch = [&](float4 (&dst)[IPT], int it) {
      const int r0 = (split + it * n_split) * kI8TileRows;
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        const int e = tid + q * (kI8ProdWarps * 32);
        const int row = e >> 4, c4 = (e & 15) * 4;
        dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < my_tiles && r0 + row < M) dst[q] = *reinterpret_cast<const float4*>(x + (size_t)(r0 + row) * 64 + c4);
      }
    };
    auto convert = [&](const float4 (&cur)[IPT], int it) {
      const int s = it & 1;
      if (it >= 2) mbar_wait(&sm.empty[s], ((it >> 1) - 1) & 1);
      unsigned char* st = sm.stage[s];
      if (!(I8_MODE & 1)) {
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
          const int e = tid + q * (kI8ProdWarps * 32);
          const int row = e >> 4, c4 = (e & 15) * 4;
          const float4 v = cur[q];
          const unsigned w0 = __float2uint_rz(v.x * 16777216.0f), w1 = __float2uint_rz(v.y * 16777216.0f);
          const unsigned w2 = __float2uint_rz(v.z * 16777216.0f), w3 = __float2uint_rz(v.w * 16777216.0f);
          const unsigned a = __byte_perm(w0, w1, 0x5140), b = __byte_perm(w0, w1, 0x7362);
          const unsigned c = __byte_perm(w2, w3, 0x5140), d = __byte_perm(w2, w3, 0x7362);
          const unsigned off = (c4 >> 4) * kI8Sbo + (row >> 3) * 128 + (row & 7) * 16 + (c4 & 15);
          *reinterpret_cast<unsigned*>(st + 0 * kI8SliceBytes + off) = __byte_perm(b, d, 0x7632);  // bits 31..24
          *reinterpret_cast<unsigned*>(st + 1 * kI8SliceBytes + off) = __byte_perm(b, d, 0x5410);  // bits 23..16
          *reinterpret_cast<unsigned*>(st + 2 * kI8SliceBytes + off) = __byte_perm(a, c, 0x7632);  // bits 15..8
          *reinterpret_cast<unsigned*>(st + 3 * kI8SliceBytes + off) = __byte_perm(a, c, 0x5410);  // bits 7..0
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (MMA)
      mbar_arrive(&sm.full[s]);
    };
    fetch(ring[0], 0);
    fetch(ring[1], 1);
    for (int it = 0; it < my_tiles; it += 3) {  // unrolled by the ring length: every ring index is a compile-time constant
      fetch(ring[2], it + 2);
      convert(ring[0], it);
      if (it + 1 < my_tiles) {
        fetch(ring[0], it + 3);
        convert(ring[1], it + 1);
      }
      if (it + 2 < my_tiles) {
        fetch(ring[1], it + 4);
        convert(ring[2], it + 2);
      }
    }
  } else if (lane == 0) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: D = s32, A = B = u8, both MN-major, N = 256, M = 128
    const unsigned idesc = (2u << 4) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      mbar_wait(&sm.full[s], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned base = smem_u32(sm.stage[s]);
#pragma unroll
      for (int j = 0; j < kI8TileRows / 32; ++j) {
        const unsigned long long bdesc = umma_desc(base + j * 512, 128, kI8Sbo);
        const unsigned long long a01 = bdesc;                                            // slices 0,1 = first 128 MN rows
        const unsigned long long a23 = umma_desc(base + 2 * kI8SliceBytes + j * 512, 128, kI8Sbo);
        const unsigned acc = (it > 0 || j > 0) ? 1u : 0u;
        if (!(I8_MODE & 2)) {
          umma_i8(tmem + 0, a01, bdesc, idesc, acc);
          umma_i8(tmem + 256, a23, bdesc, idesc, acc);
        }
      }
      // frees the stage when the MMAs above have consumed it (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.done)) : "memory");
  }

  // ---------------- epilogue: D (128 x 512 int32 in TMEM) -> G (64 x 64 f64) ----------------
  double* gpart = reinterpret_cast<double*>(sm.stage[0]);  // [128][64] doubles = 64 KB (spans both stages)
  if (warp < 4) {
    if (my_tiles > 0) {
      mbar_wait(&sm.done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int r = warp * 32 + lane;      // accumulator row: slice a = r / 64 (+2 for D2), G row n = r % 64
      const int a1 = r >> 6;
      const unsigned lane_addr = tmem + ((unsigned)(warp * 32) << 16);
      for (int c0 = 0; c0 < 64; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // D1 then D2
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            unsigned v[16];
            tmem_ld16(lane_addr + half * 256 + b * 64 + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const double scale = exp2(-8.0 * (double)(a1 + 2 * half + b));
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fma((double)v[j], scale, acc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) gpart[r * 64 + c0 + j] = acc[j];
      }
    } else {
      for (int j = 0; j < 64; ++j) gpart[(warp * 32 + lane) * 64 + j] = 0.0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  double* g = Gout + ((size_t)mat * n_split + split) * 4096;
  for (int e = tid; e < 4096; e += kI8Threads) {
    const int n = e >> 6, m = e & 63;
    const int lo = n < m ? n : m, hi = n < m ? m : n;  // symmetric output from the upper triangle
    g[e] = gpart[lo * 64 + hi] + gpart[(lo + 64) * 64 + hi];
  }
  __syncthreads();
  if (warp == kI8ProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace lrfb

#endif  // LRFB_SIM

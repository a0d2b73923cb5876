// Front end fused with the luma Gram (sm_100a only): one pass over the uint8 image produces the three patch matrices
// (Y, Cb, Cr as f32, the same bits as frontend8_fused_kernel) AND G_y = X_y^T X_y on the int8 tensor cores, so the
// 4 B/pixel that the separate Gram kernel reads back from HBM are never read.  Geometry: the fused front end's
// (uint8, 8x8 patches, H % 16 == 0, no padding) with W % 256 == 0; one CTA per image.
//
// A tile is the front end's 16-row x 256-pixel block: 2 x 32 luma patches (K = 64 rows of X_y: two k-steps of the
// Gram MMAs) and 16 chroma patches per plane.  Producer thread (cy, wp) owns the 2-row x 8-pixel strip of the three
// channels, exactly as in frontend8_fused_kernel; it stores its 2 x 8 luma values and 4 + 4 chroma values to X and
// the Q8.24 bytes of the luma values into the MN-major core-matrix staging of gram64_i8_kernel (4 stages of 64 rows).
// Two producer groups of 8 warps work on alternate tiles (one CTA per SM because of the 512 TMEM columns: the second
// group is what hides the load and conversion latency of the first); the last warp issues the MMAs; warps 0-3 run the
// shared epilogue.
#pragma once
#include "frontend.cuh"
#include "gram_i8.cuh"

#ifndef LRFB_SIM

namespace lrfb {

constexpr int kFgTileRows = 64;                               // luma patches per tile
constexpr int kFgSbo = (kFgTileRows / 8) * 128 + 32;          // MN-core stride (8 K-cores of 128 B + bank padding)
constexpr int kFgSliceBytes = 4 * kFgSbo;
constexpr int kFgStageBytes = 4 * kFgSliceBytes;              // 16 896 B
constexpr int kFgStages = 4;
constexpr int kFgGroups = 2;                                  // producer groups of 256 threads, one tile each at a time
constexpr int kFgProdWarps = 8 * kFgGroups;
constexpr int kFgThreads = (kFgProdWarps + 1) * 32;           // + the MMA warp

struct FrontGramSmem {
  unsigned char stage[kFgStages][kFgStageBytes];  // >= 64 KB: reused as the 128 x 64 f64 partials of the epilogue
  // per producer group: the tile's patches staged for fully coalesced stores to X (as in frontend8_fused_kernel;
  // 16-byte stores straight from the strips cost one L1 transaction per lane and left the kernel LSU-bound)
  float xl[kFgGroups][2 * 32 * kTilePatchStride];
  float xcb[kFgGroups][16 * kTilePatchStride];
  float xcr[kFgGroups][16 * kTilePatchStride];
  unsigned long long full[kFgStages], empty[kFgStages], done;
  unsigned tmem_base;
};
static_assert(kFgStages * kFgStageBytes >= 128 * 64 * 8, "epilogue scratch");

// D (128 x 512 int32 in TMEM: [S0;S1] and [S2;S3] against all four slices) -> G (64 x 64 f64), shared with
// gram64_i8_kernel's arithmetic: sum_{a,b} 2^(-8(a+b)) S_a^T S_b in f64.
__device__ __forceinline__ void gram_i8_epilogue(double* gpart, unsigned long long* done_bar, unsigned tmem, bool any,
                                                 int tid, int warp, int lane, int n_threads, double* g) {
  if (warp < 4) {
    if (any) {
      mbar_wait(done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int r = warp * 32 + lane;
      const int a1 = r >> 6;
      const unsigned lane_addr = tmem + ((unsigned)(warp * 32) << 16);
      for (int c0 = 0; c0 < 64; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
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
  for (int e = tid; e < 4096; e += n_threads) {
    const int n = e >> 6, m = e & 63;
    const int lo = n < m ? n : m, hi = n < m ? m : n;  // symmetric output from the upper triangle
    g[e] = gpart[lo * 64 + hi] + gpart[(lo + 64) * 64 + hi];
  }
  __syncthreads();
}

// grid = images (strided).  xy / xcb / xcr: the three patch matrices [n_img][rows][64]; Gy: [n_img][64][64] f64.
__global__ void __launch_bounds__(kFgThreads, 1)
frontgram8_kernel(const unsigned char* __restrict__ images, float* __restrict__ xy, float* __restrict__ xcb,
                  float* __restrict__ xcr, double* __restrict__ Gy, FrontParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FrontGramSmem& sm = *reinterpret_cast<FrontGramSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PlaneGeom gl = P.g[0], gc = P.g[1];
  const size_t hw = (size_t)P.H * P.W;
  const int nbl = gl.nbw, nbc = gc.nbw;    // W/8 and W/16 patches per patch row
  const int tiles_x = P.W / 256, n_tiles = (P.H / 16) * tiles_x;

  if (tid == 0) {
    for (int s = 0; s < kFgStages; ++s) mbar_init(&sm.full[s], 256), mbar_init(&sm.empty[s], 1);
    mbar_init(&sm.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFgProdWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = sm.tmem_base;
  int tiles_done = 0;  // running tile count of this CTA: stage = count % kFgStages, phase from count / kFgStages

  for (int im = blockIdx.x; im < P.n_img; im += gridDim.x) {
    const unsigned char* img = images + (size_t)im * 3 * hw;
    if (warp < kFgProdWarps) {
      // ---------------- producers: group `grp` takes tiles grp, grp + 2, ... ----------------
      const int grp = tid >> 8, wp = tid & 31, cy = (tid >> 5) & 7;
      float* oy = xy + (size_t)im * gl.rows * 64;
      float* ocb = xcb + (size_t)im * gc.rows * 64;
      float* ocr = xcr + (size_t)im * gc.rows * 64;
      uint2 nr[2], ng[2], nb[2];
      auto fetch = [&](int t) {
        if (t >= n_tiles) return;
        const int y0 = (t / tiles_x) * 16, wp0 = (t % tiles_x) * 32;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const size_t off = (size_t)(y0 + 2 * cy + dy) * P.W + (size_t)(wp0 + wp) * 8;
          nr[dy] = *reinterpret_cast<const uint2*>(img + off);
          ng[dy] = *reinterpret_cast<const uint2*>(img + hw + off);
          nb[dy] = *reinterpret_cast<const uint2*>(img + 2 * hw + off);
        }
      };
      fetch(grp);
      for (int t = grp; t < n_tiles; t += kFgGroups) {
        const int y0 = (t / tiles_x) * 16, wp0 = (t % tiles_x) * 32;
        uint2 cr_[2] = {nr[0], nr[1]}, cg_[2] = {ng[0], ng[1]}, cb_[2] = {nb[0], nb[1]};
        fetch(t + kFgGroups);  // this group's next tile: its loads fly while this one is converted
        const int cnt = tiles_done + t, s = cnt % kFgStages;
        if (cnt >= kFgStages) mbar_wait(&sm.empty[s], ((cnt / kFgStages) - 1) & 1);
        unsigned char* st = sm.stage[s];
        float* s_lum = sm.xl[grp];
        float* s_cb = sm.xcb[grp];
        float* s_cr = sm.xcr[grp];
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");  // the group's previous write-out is done with s_*
        float sb[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        unsigned sl[4][4];  // [slice][word]: the 16 bytes of this strip's core-matrix row per slice
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ly = 2 * cy + dy;
          const unsigned rw[2] = {cr_[dy].x, cr_[dy].y}, gw[2] = {cg_[dy].x, cg_[dy].y}, bw[2] = {cb_[dy].x, cb_[dy].y};
          float lum[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int px = 2 * j + q;
              const float fr = byte_of(rw[px >> 2], px & 3), fg = byte_of(gw[px >> 2], px & 3),
                          fb = byte_of(bw[px >> 2], px & 3);
              lum[px] = ycc_from(fr, fg, fb, 0);
              sb[j] = __fadd_rn(sb[j], ycc_from(fr, fg, fb, 1));
              sr[j] = __fadd_rn(sr[j], ycc_from(fr, fg, fb, 2));
            }
          }
          // X_y staging: tile patch ((ly >> 3), wp), elements (ly & 7) * 8 .. + 7
          float* d = &s_lum[((ly >> 3) * 32 + wp) * kTilePatchStride + (ly & 7) * 8];
          *reinterpret_cast<float4*>(d) = make_float4(lum[0], lum[1], lum[2], lum[3]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(lum[4], lum[5], lum[6], lum[7]);
          // Q8.24 byte slices of the 8 values: tile row m, columns n = (ly & 7) * 8 .. + 7.  The strip's two image rows
          // are the two halves of one 16-byte core-matrix row, stored together below (conflict-free 16-byte stores).
          unsigned w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = __float2uint_rz(lum[k] * 16777216.0f);
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // values 4h .. 4h+3 packed into one 32-bit word per slice
            const unsigned a = __byte_perm(w[4 * h], w[4 * h + 1], 0x5140), b = __byte_perm(w[4 * h], w[4 * h + 1], 0x7362);
            const unsigned c = __byte_perm(w[4 * h + 2], w[4 * h + 3], 0x5140), e = __byte_perm(w[4 * h + 2], w[4 * h + 3], 0x7362);
            sl[0][2 * dy + h] = __byte_perm(b, e, 0x7632);  // bits 31..24
            sl[1][2 * dy + h] = __byte_perm(b, e, 0x5410);  // bits 23..16
            sl[2][2 * dy + h] = __byte_perm(a, c, 0x7632);  // bits 15..8
            sl[3][2 * dy + h] = __byte_perm(a, c, 0x5410);  // bits 7..0
          }
        }
        {
          const int ly = 2 * cy, m = (ly >> 3) * 32 + wp, n0 = (ly & 7) * 8;  // n0 is a multiple of 16
          const unsigned off = (n0 >> 4) * kFgSbo + (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
          for (int a = 0; a < 4; ++a)
            *reinterpret_cast<uint4*>(st + a * kFgSliceBytes + off) = make_uint4(sl[a][0], sl[a][1], sl[a][2], sl[a][3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (MMA)
        mbar_arrive(&sm.full[s]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // (s / 2) / 2: exact power-of-two scalings
          sb[j] = __fmul_rn(__fmul_rn(sb[j], 0.5f), 0.5f);
          sr[j] = __fmul_rn(__fmul_rn(sr[j], 0.5f), 0.5f);
        }
        const int cs = (wp >> 1) * kTilePatchStride + cy * 8 + (wp & 1) * 4;
        *reinterpret_cast<float4*>(&s_cb[cs]) = make_float4(sb[0], sb[1], sb[2], sb[3]);
        *reinterpret_cast<float4*>(&s_cr[cs]) = make_float4(sr[0], sr[1], sr[2], sr[3]);
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");  // the tile is staged
        // write-out: 2 runs of 32 luma patches, 1 run of 16 patches per chroma plane, all contiguous in X
        const int lt = tid & 255;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int i = lt + q4 * 256, pr = i >> 9, patch = (i >> 4) & 31, q = i & 15;
          const float4 v = *reinterpret_cast<const float4*>(&s_lum[(pr * 32 + patch) * kTilePatchStride + q * 4]);
          *reinterpret_cast<float4*>(oy + ((size_t)((y0 >> 3) + pr) * nbl + wp0 + patch) * 64 + q * 4) = v;
        }
        const int patch = lt >> 4, q = lt & 15;
        const size_t co = ((size_t)(y0 >> 4) * nbc + (wp0 >> 1) + patch) * 64 + q * 4;
        *reinterpret_cast<float4*>(ocb + co) = *reinterpret_cast<const float4*>(&s_cb[patch * kTilePatchStride + q * 4]);
        *reinterpret_cast<float4*>(ocr + co) = *reinterpret_cast<const float4*>(&s_cr[patch * kTilePatchStride + q * 4]);
      }
    } else if (lane == 0) {
      // ---------------- MMA issuer: D = s32, A = B = u8, both MN-major, N = 256, M = 128 ----------------
      const unsigned idesc = (2u << 4) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      for (int t = 0; t < n_tiles; ++t) {
        const int cnt = tiles_done + t, s = cnt % kFgStages;
        mbar_wait(&sm.full[s], (cnt / kFgStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned base = smem_u32(sm.stage[s]);
#pragma unroll
        for (int j = 0; j < kFgTileRows / 32; ++j) {
          const unsigned long long bdesc = umma_desc(base + j * 512, 128, kFgSbo);
          const unsigned long long a23 = umma_desc(base + 2 * kFgSliceBytes + j * 512, 128, kFgSbo);
          const unsigned acc = (t > 0 || j > 0) ? 1u : 0u;
          umma_i8(tmem + 0, bdesc, bdesc, idesc, acc);
          umma_i8(tmem + 256, a23, bdesc, idesc, acc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.done)) : "memory");
    }
    tiles_done += n_tiles;
    // every stage has been consumed once `done` fires (it follows the last tile's MMAs), so the staging memory is free
    gram_i8_epilogue(reinterpret_cast<double*>(sm.stage[0]), &sm.done, tmem, true, tid, warp, lane, kFgThreads,
                     Gy + (size_t)im * 4096);
    break;  // one image per CTA: the mbarrier phases of `done` / `empty` are not re-armed for a second one
  }
  if (warp == kFgProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace lrfb

#endif  // LRFB_SIM

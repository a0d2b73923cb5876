// Resident QMF sweeps with the V-phase on the 5th-generation tensor cores (sm_100a only).
//
// Same structure as bcd_resident.cuh (cluster of 1/2/4/8 CTAs per matrix, 768 rows of X per CTA resident in
// shared memory as f32 for the bit-exact A-phase, DSMEM exchange of the per-CTA partial sums), but the K = M
// reduction S = X^T U — the half of every sweep whose summation order the reference leaves opaque
// and which costs as much as the A-phase on the FFMA pipe — is computed EXACTLY as an integer product:
//   * once per matrix each CTA converts its X rows to Q8.24 fixed point (x * 2^24, exact for x >= 0.5, entries
//     are in [0, 256)) and parks the four byte planes S_0..S_3 in TENSOR MEMORY as the A operand
//     ([S_0;S_1] and [S_2;S_3], 128 lanes x 192 columns each: lane = (slice, column n), 4 rows per 32-bit cell);
//   * every sweep the freshly projected U rows (integers in [-128,127]) are written as int8 in the K-major
//     core-matrix layout to shared memory (6 KB), and one thread issues 2 x 24 tcgen05.mma kind::i8
//     (M = 128, N = 8, K = 32; A from TMEM, B = U from shared memory) into two 128 x 8 int32 accumulators;
//   * S[n][r] = sum_a 2^(-8a) D_a[n][r] is exact in f64 (<= 52 bits), so X^T U carries no rounding at all
//     before its single conversion to f32 — closer to the reference's intent than any f32 summation order.
// The FFMA pipe and the shared-memory pipe are left to the A-phase alone.
#pragma once
#include "bcd_resident.cuh"
#include "gram_i8.cuh"

#ifndef LRFB_SIM

namespace lrfb {

constexpr int kTcRows = 768;  // largest row slice per CTA (1 CTA per SM); the 384-row shape runs 2 CTAs per SM

template <int R, int ROWS, int NT>
struct TcSmem {
  static constexpr int N = 64;
  float x[ROWS * N];                    // swizzled f32 rows (A-phase)
  unsigned char ub[ROWS * 8];           // B operand: U as int8, K-major cores: (m/16)*128 + r*16 + m%16
  float v[N * R];
  float b[R * R];
  float b2[R * R];
  float a2[N * R];
  float s0inv[4];
  int gred[(NT / 32) * R * R];
  double comb[128 * 4];                 // per accumulator lane: slices already combined
  double part[2][N * R + R * R];
  unsigned long long mma_done, clear_done;
  unsigned tmem_base;
};

__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void umma_i8_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long b, unsigned idesc,
                                           unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int R, int ROWS, int NT>
__global__ void __launch_bounds__(NT, (ROWS <= 384 ? 2 : 1))  // 2 CTAs/SM: the tail of one overlaps the other's compute
bcd_tc_kernel(BcdBatch P, int cluster_size, int rows_per_cta) {
  constexpr int N = 64, RT = ROWS / NT, NW = NT / 32;
  constexpr int kTcRows = ROWS;
  constexpr int kTcColsA = ROWS / 4;                        // TMEM columns per A block (4 rows per cell)
  constexpr int kTcD1 = 2 * kTcColsA + 16, kTcD2 = kTcD1 + 16;  // accumulator columns
  constexpr int kTmemCols = ROWS > 384 ? 512 : 256;
  static_assert(ROWS % NT == 0 && ROWS % 64 == 0 && NW >= 4 && kTcD2 + 8 <= kTmemCols, "shape");
  using S = TcSmem<R, ROWS, NT>;
  LRFB_DYN_SMEM(smem_raw);
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = P.M;
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / cluster_size;
  const int n_clusters = gridDim.x / cluster_size;
  const bool t2_native_u = bmm_native(R - 1, M, 1);
  constexpr bool t2_native_v = (long long)(R - 1) * N < 400;
  const bool from_a = P.s0 != nullptr;
  const int row0 = crank * rows_per_cta;
  const int rows_here = max(0, min(rows_per_cta, M - row0));
  int pbuf = 0;
  unsigned mma_phase = 0;

  if (tid == 0) {
    mbar_init(&sm.mma_done, NW);  // one tcgen05.commit per warp and sweep
    mbar_init(&sm.clear_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < kTcRows * 8; i += NT) sm.ub[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = sm.tmem_base;
  // D = s32, A = u8 (TMEM), B = s8 K-major, N = 8, M = 128
  const unsigned idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
  // The accumulators are never cleared between sweeps: integer accumulation is exact and order-free, so every
  // warp issues the MMAs of its own rows as soon as they are projected, and the epilogue takes differences
  // (mod 2^32) against the previous sweep's totals.  One product against the all-zero B zeroes them here.
  int prev1[8], prev2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) prev1[i] = prev2[i] = 0;
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned long long bd0 = umma_desc(smem_u32(sm.ub), 128, 128);
    umma_i8_ts(tmem + kTcD1, tmem, bd0, idesc, 0);
    umma_i8_ts(tmem + kTcD2, tmem + kTcColsA, bd0, idesc, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.clear_done)) : "memory");
  }
  mbar_wait(&sm.clear_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  for (int mat = cluster_id; mat < P.n_mat; mat += n_clusters) {
    const float* X = P.X + (size_t)mat * P.x_stride;
    float* V = P.V + (size_t)mat * N * R;

    // ---- load this CTA's slice of X once (swizzled f32) and V ----
    __syncthreads();
    for (int c = tid; c < kTcRows * (N / 4); c += NT) {
      const int row = c >> 4, ch = c & 15;
      float* dst = &sm.x[row * N + ((ch ^ (row & 7)) << 2)];
      if (row < rows_here) cp_async16(dst, X + (size_t)(row0 + row) * N + ch * 4);
      else dst[0] = dst[1] = dst[2] = dst[3] = 0.0f;
    }
    cp_async_commit();
    for (int i = tid; i < N * R; i += NT) sm.v[i] = V[i];
    if (tid < R) {
      float inv = 0.0f;
      if (from_a) {
        const float sv = P.s0[(size_t)mat * R + tid];
        inv = sv > 0.0f ? __fdiv_rn(1.0f, sv) : 0.0f;
      }
      sm.s0inv[tid] = inv;
    }
    const float* Uinit = P.U + (size_t)mat * M * R + (size_t)row0 * R;
    cp_async_wait<0>();
    __syncthreads();
    gram_small<N, R>(sm.v, sm.b, tid);

    // ---- Q8.24 byte planes of X into tensor memory (A operand of the V-phase MMAs) ----
    {
      const int l = (warp & 3) * 32 + lane;     // accumulator / operand lane of this thread's TMEM quarter
      const int a_lo = l >> 6, n = l & 63;      // slice a_lo in block 0, a_lo + 2 in block 1
      const int sh0 = 24 - 8 * a_lo, sh1 = 8 - 8 * a_lo;
      const unsigned lane_addr = tmem + ((unsigned)((warp & 3) * 32) << 16);
      const int sharers = (NW - (warp & 3) + 3) / 4;  // warps that own this TMEM lane quarter
      for (int ch = warp >> 2; ch < kTcRows / 64; ch += sharers) {  // 64 rows (16 cells) per store
        unsigned w0[16], w1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          unsigned p0 = 0, p1 = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int m = ch * 64 + c * 4 + t;
            const float xv = sm.x[m * N + (((n >> 2) ^ (m & 7)) << 2) + (n & 3)];
            const unsigned fx = __float2uint_rz(xv * 16777216.0f);
            p0 |= ((fx >> sh0) & 0xffu) << (8 * t);
            p1 |= ((fx >> sh1) & 0xffu) << (8 * t);
          }
          w0[c] = p0, w1[c] = p1;
        }
        tmem_st16(lane_addr + ch * 16, w0);
        tmem_st16(lane_addr + kTcColsA + ch * 16, w1);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // the first kReg A-phase rows of this thread stay in registers for all sweeps
    constexpr int kReg = (NT == 128) ? 2 : (R == 4 ? 1 : 0);
    constexpr int kRegN = kReg ? N : 1;
    float xr[kReg ? kReg : 1][kRegN];
#pragma unroll
    for (int i = 0; i < kReg; ++i) {
      const int row = tid + i * NT;
#pragma unroll
      for (int k4 = 0; k4 < N / 4; ++k4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&sm.x[row * N + ((k4 ^ (row & 7)) << 2)]);
        xr[i][(4 * k4 + 0) % kRegN] = t4.x, xr[i][(4 * k4 + 1) % kRegN] = t4.y;
        xr[i][(4 * k4 + 2) % kRegN] = t4.z, xr[i][(4 * k4 + 3) % kRegN] = t4.w;
      }
    }
    float uown[RT][R];  // this thread's U rows live in registers across the sweeps
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    for (int it = 0; it < P.num_iters; ++it) {
      // ---------------- A-phase + Gauss–Seidel (identical arithmetic to bcd_resident_kernel) ----------------
      int gacc[R * (R + 1) / 2];
#pragma unroll
      for (int i = 0; i < R * (R + 1) / 2; ++i) gacc[i] = 0;
      {
        float acc[RT][R];
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
          for (int r = 0; r < R; ++r) acc[i][r] = 0.0f;
#pragma unroll
        for (int k4 = 0; k4 < N / 4; ++k4) {
          float vk[4][R];
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) vk[k][r] = sm.v[(k4 * 4 + k) * R + r];
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            const int row = tid + i * NT;
            float4 xv;
            if (i < kReg) {
              constexpr int z = 0;
              const int ii = i < kReg ? i : z;
              xv = make_float4(xr[ii][(4 * k4 + 0) % kRegN], xr[ii][(4 * k4 + 1) % kRegN],
                               xr[ii][(4 * k4 + 2) % kRegN], xr[ii][(4 * k4 + 3) % kRegN]);
            } else {
              xv = *reinterpret_cast<const float4*>(&sm.x[row * N + ((k4 ^ (row & 7)) << 2)]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
              float a = acc[i][r];
              a = __fmaf_rn(xv.x, vk[0][r], a);
              a = __fmaf_rn(xv.y, vk[1][r], a);
              a = __fmaf_rn(xv.z, vk[2][r], a);
              a = __fmaf_rn(xv.w, vk[3][r], a);
              acc[i][r] = a;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const int row = tid + i * NT;
          const bool ok = row < rows_here;
          float f[R];
          if (it == 0) {
            if (from_a) {
#pragma unroll
              for (int r = 0; r < R; ++r) f[r] = sm.s0inv[r] == 0.0f ? 0.0f : __fmul_rn(acc[i][r], sm.s0inv[r]);
            } else {
#pragma unroll
              for (int r = 0; r < R; ++r) f[r] = ok ? Uinit[row * R + r] : 0.0f;
            }
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) f[r] = uown[i][r];
          }
          gs_row<R>(f, acc[i], sm.b, t2_native_u, P.lo, P.hi);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (!ok) f[r] = 0.0f;
            uown[i][r] = f[r];
            sm.ub[(row >> 4) * 128 + r * 16 + (row & 15)] = (unsigned char)(signed char)(int)f[r];
          }
          if (ok) {
            int idx = 0;
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
              for (int r = j; r < R; ++r) gacc[idx++] += (int)f[j] * (int)f[r];
          }
        }
      }
      // U^T U partial of this warp: exact integers, one REDUX per entry
      {
        int idx = 0;
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
          for (int r = j; r < R; ++r) {
            const int g = __reduce_add_sync(0xffffffffu, gacc[idx++]);
            if (lane == 0) sm.gred[warp * R * R + j * R + r] = g, sm.gred[warp * R * R + r * R + j] = g;
          }
      }
      // ---------------- V-phase on the tensor core: D_a += S_a^T U over this warp's 2 x 32 rows ----------------
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // int8 U rows -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned bbase = smem_u32(sm.ub);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const int j = warp + i * NW;  // k-chunk of 32 rows: rows 32*warp.. (i = 0) and 384 + 32*warp.. (i = 1)
          const unsigned long long bd = umma_desc(bbase + j * 256, 128, 128);
          umma_i8_ts(tmem + kTcD1, tmem + j * 8, bd, idesc, 1);
          umma_i8_ts(tmem + kTcD2, tmem + kTcColsA + j * 8, bd, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.mma_done)) : "memory");
      }
      if (warp < 4) {
        mbar_wait(&sm.mma_done, mma_phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int l = warp * 32 + lane, a_lo = l >> 6;
        unsigned d1[8], d2[8];
        tmem_ld8(tmem + ((unsigned)(warp * 32) << 16) + kTcD1, d1);
        tmem_ld8(tmem + ((unsigned)(warp * 32) << 16) + kTcD2, d2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const double s1 = a_lo ? 0.00390625 : 1.0;                       // 2^(-8 a)
        const double s2 = a_lo ? 5.9604644775390625e-08 : 1.52587890625e-05;  // 2^(-8 (a+2))
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e1 = (int)d1[r] - prev1[r], e2 = (int)d2[r] - prev2[r];  // this sweep's sums (exact mod 2^32)
          prev1[r] = (int)d1[r], prev2[r] = (int)d2[r];
          sm.comb[l * 4 + r] = (double)e1 * s1 + (double)e2 * s2;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      mma_phase ^= 1;
      __syncthreads();
      for (int e = tid; e < N * R + R * R; e += NT) {
        if (e < N * R) {
          const int n = e / R, r = e - n * R;
          sm.part[pbuf][e] = sm.comb[n * 4 + r] + sm.comb[(n + 64) * 4 + r];  // exact: <= 52 significant bits
        } else {
          int g = 0;
#pragma unroll
          for (int w2 = 0; w2 < NW; ++w2) g += sm.gred[w2 * R * R + e - N * R];
          sm.part[pbuf][e] = (double)g;
        }
      }

      // ---------------- exchange partials across the cluster, every CTA sums in rank order ----------------
      if (cluster_size > 1) {
        cluster.sync();
        for (int e = tid; e < N * R + R * R; e += NT) {
          double s = 0.0;
          for (int cr = 0; cr < cluster_size; ++cr) {
            const double* remote = cluster.map_shared_rank(&sm.part[pbuf][0], cr);
            s += remote[e];
          }
          if (e < N * R) sm.a2[e] = (float)s;
          else sm.b2[e - N * R] = (float)s;
        }
      } else {
        __syncthreads();
        for (int e = tid; e < N * R + R * R; e += NT) {
          if (e < N * R) sm.a2[e] = (float)sm.part[pbuf][e];
          else sm.b2[e - N * R] = (float)sm.part[pbuf][e];
        }
      }
      pbuf ^= 1;
      __syncthreads();

      // ---------------- V update (identical in every CTA of the cluster) and B = V^T V ----------------
      for (int n = tid; n < N; n += NT) {
        float f[R], A[R];
#pragma unroll
        for (int r = 0; r < R; ++r) f[r] = sm.v[n * R + r], A[r] = sm.a2[n * R + r];
        gs_row<R>(f, A, sm.b2, t2_native_v, P.lo, P.hi);
#pragma unroll
        for (int r = 0; r < R; ++r) sm.v[n * R + r] = f[r];
      }
      __syncthreads();
      {
        for (int e = warp; e < R * R; e += NW) {  // V is integer-valued: every order gives the same exact result
          const int j = e / R, r = e - j * R;
          float p = __fmaf_rn(sm.v[lane * R + j], sm.v[lane * R + r],
                              __fmul_rn(sm.v[(lane + 32) * R + j], sm.v[(lane + 32) * R + r]));
          for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
          if (lane == 0) sm.b[e] = p;
        }
      }
      __syncthreads();
    }

    // ---- write the factors of this CTA's rows (and V once per cluster) ----
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const int row = tid + i * NT;
      if (row < rows_here) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (P.U) P.U[(size_t)mat * M * R + (size_t)(row0 + row) * R + r] = uown[i][r];
          if (P.Uq) P.Uq[(size_t)mat * P.uq_stride + (size_t)r * M + row0 + row] = (int8_t)(int)uown[i][r];
        }
      }
    }
    if (crank == 0) {
      for (int i = tid; i < N * R; i += NT) {
        V[i] = sm.v[i];
        if (P.Vq) P.Vq[(size_t)mat * P.vq_stride + (size_t)(i % R) * N + i / R] = (int8_t)(int)sm.v[i];
      }
    }
  }
  if (cluster_size > 1) cluster.sync();  // nobody leaves while its partials may still be read
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace lrfb

#endif  // LRFB_SIM

// Shared-memory-resident variant of the QMF block-coordinate-descent sweeps (same arithmetic as bcd.cuh,
// lrf/factorization/qmf.py:93-139, :191-214) for matrices of up to 8 x 768 rows with N = 64.
//
// A thread-block CLUSTER of C = 1/2/4/8 CTAs owns one matrix for all sweeps.  Each CTA loads its row
// slice of X (<= 768 rows = 192 KB) into shared memory ONCE and keeps it there: after the first pass the
// sweeps touch HBM only for the final factors.  Per sweep each CTA does the A-phase / Gauss–Seidel /
// V-phase on its rows, the per-CTA partial sums of X^T U (f64) and U^T U (exact integers) are exchanged
// through distributed shared memory, summed by every CTA in the same fixed rank order, and every CTA
// performs the identical 64 x R update of V redundantly — one cluster barrier per sweep, partial buffers
// double-buffered.  Shared memory is XOR-swizzled at 16-byte granularity (chunk ^= row & 7) so both the
// row-per-thread (A-phase) and the row-per-half-warp (V-phase) 128-bit reads are bank-conflict free
// without padding.
#pragma once
#include "bcd.cuh"

#ifndef LRFB_SIM
#include <cooperative_groups.h>
#endif

namespace lrfb {


// U rows are kept as bf16 (exact for the integer range [-128, 127]; f32 <-> bf16 is a 16-bit shift).
__device__ __forceinline__ unsigned short bf16_of(float f) { return (unsigned short)(__float_as_uint(f) >> 16); }
__device__ __forceinline__ float bf16_to(unsigned short h) { return __uint_as_float((unsigned)h << 16); }

template <int R, int ROWS, int NT>
struct ResSmem {
  static constexpr int N = 64;
  static constexpr int NW = NT / 32;
  float x[ROWS * N];              // swizzled
  unsigned short u[ROWS * R];     // current U rows of this CTA (bf16)
  float v[N * R];
  float b[R * R];
  float b2[R * R];
  float a2[N * R];
  float s0inv[4];
  float red[NW * N * R];          // per-warp f32 chunk sums of X^T U, summed in f64 in warp order
  int gred[NW * R * R];
  double part[2][N * R + R * R];  // this CTA's partial S and U^T U, read by the whole cluster
};

__device__ __forceinline__ void cluster_barrier() {
#ifndef LRFB_SIM
  cooperative_groups::this_cluster().sync();
#else
  __syncthreads();
#endif
}

template <int R, int ROWS, int NT>
__global__ void __launch_bounds__(NT, (ROWS <= 384 ? 2 : 1))
bcd_resident_kernel(BcdBatch P, int cluster_size, int rows_per_cta) {
  constexpr int N = 64;
  constexpr int kResRows = ROWS;
  constexpr int RT = kResRows / NT;  // rows per thread in the A-phase
  constexpr int NG = NT / 16;        // half-warp row groups in the V-phase
  constexpr int NW = NT / 32;
  constexpr int kRegRows = (R == 4 && ROWS == 768 && NT == 384) ? 1 : 0;  // register-resident A-phase row
  static_assert(kResRows % NT == 0, "thread shape");
  using S = ResSmem<R, ROWS, NT>;
  LRFB_DYN_SMEM(smem_raw);
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int tid = threadIdx.x;
  const int M = P.M;
#ifndef LRFB_SIM
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int crank = (int)cluster.block_rank();
#else
  const int crank = 0;
#endif
  const int cluster_id = blockIdx.x / cluster_size;
  const int n_clusters = gridDim.x / cluster_size;
  const bool t2_native_u = bmm_native(R - 1, M, 1);
  constexpr bool t2_native_v = (long long)(R - 1) * N < 400;
  const bool from_a = P.s0 != nullptr;
  const int row0 = crank * rows_per_cta;                          // first row of this CTA
  const int rows_here = max(0, min(rows_per_cta, M - row0));
  int pbuf = 0;

  for (int mat = cluster_id; mat < P.n_mat; mat += n_clusters) {
    const float* X = P.X + (size_t)mat * P.x_stride;
    float* V = P.V + (size_t)mat * N * R;

    // ---- load this CTA's slice of X once (swizzled), V, and the injected U init if any ----
    __syncthreads();
    for (int c = tid; c < kResRows * (N / 4); c += NT) {
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
    const float* Uinit = P.U + (size_t)mat * M * R + (size_t)row0 * R;  // read in sweep 1 only when injected
    cp_async_wait<0>();
    __syncthreads();
    gram_small<N, R>(sm.v, sm.b, tid);
    // the first of this thread's A-phase rows stays in registers for all sweeps (no shared-memory reads
    // for it in the A-phase; the V-phase still reads it from shared memory)
    float xr[kRegRows ? N : 1];
    if (kRegRows) {
#pragma unroll
      for (int k4 = 0; k4 < N / 4; ++k4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&sm.x[tid * N + ((k4 ^ (tid & 7)) << 2)]);
        xr[(4 * k4 + 0) % (kRegRows ? N : 1)] = t4.x, xr[(4 * k4 + 1) % (kRegRows ? N : 1)] = t4.y;
        xr[(4 * k4 + 2) % (kRegRows ? N : 1)] = t4.z, xr[(4 * k4 + 3) % (kRegRows ? N : 1)] = t4.w;
      }
    }
    __syncthreads();

    for (int it = 0; it < P.num_iters; ++it) {
      // ---------------- A-phase + Gauss–Seidel ----------------
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
            if (kRegRows && i == 0) {
              xv = make_float4(xr[(4 * k4 + 0) % (kRegRows ? N : 1)], xr[(4 * k4 + 1) % (kRegRows ? N : 1)],
                               xr[(4 * k4 + 2) % (kRegRows ? N : 1)], xr[(4 * k4 + 3) % (kRegRows ? N : 1)]);
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
          float f[R];
          const bool ok = row < rows_here;
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
            for (int r = 0; r < R; ++r) f[r] = bf16_to(sm.u[row * R + r]);
          }
          gs_row_auto<R>(f, acc[i], sm.b, P.x_u8_range != 0, t2_native_u, P.lo, P.hi);
#pragma unroll
          for (int r = 0; r < R; ++r) sm.u[row * R + r] = bf16_of(ok ? f[r] : 0.0f);
          if (ok) {
            int idx = 0;
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
              for (int r = j; r < R; ++r) gacc[idx++] += (int)f[j] * (int)f[r];
          }
        }
      }
      __syncwarp();  // the V-phase of a warp only touches the rows that same warp just updated

      // ---------------- V-phase: S[n][r] += X[m][n] * U[m][r] ----------------
      // Half-warp h of warp w takes the 32*RT/2 rows w*32.. (+ h*NT ...) that warp w owns in the A-phase, so
      // no CTA barrier separates the phases and warps drift apart (A-phase FMA/LDS overlaps V-phase of
      // others).  f32 within the <=64-row chunk of a warp, chunk sums combined in f64 below.
      {
        const int ln = tid & 15, w = tid >> 5, half = (tid >> 4) & 1;
        constexpr int HR = 16 * RT;  // rows per half warp: the warp's 32*RT A-phase rows split in two
        float sacc[4][R];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < R; ++r) sacc[c][r] = 0.0f;
#pragma unroll 16
        for (int q = 0; q < HR; ++q) {
          const int t = half * HR + q;  // index into the warp's row list
          const int row = w * 32 + (t & 31) + (t >> 5) * NT;
          const float4 xv = *reinterpret_cast<const float4*>(&sm.x[row * N + ((ln ^ (row & 7)) << 2)]);
          float u[R];
          if (R == 4) {
            const uint2 w2 = *reinterpret_cast<const uint2*>(&sm.u[row * 4]);
            u[0] = __uint_as_float(w2.x << 16), u[1 % R] = __uint_as_float(w2.x & 0xffff0000u);
            u[2 % R] = __uint_as_float(w2.y << 16), u[3 % R] = __uint_as_float(w2.y & 0xffff0000u);
          } else if (R == 2) {
            const unsigned w1 = *reinterpret_cast<const unsigned*>(&sm.u[row * 2]);
            u[0] = __uint_as_float(w1 << 16), u[1 % R] = __uint_as_float(w1 & 0xffff0000u);
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = bf16_to(sm.u[row * R + r]);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            sacc[0][r] = __fmaf_rn(xv.x, u[r], sacc[0][r]);
            sacc[1][r] = __fmaf_rn(xv.y, u[r], sacc[1][r]);
            sacc[2][r] = __fmaf_rn(xv.z, u[r], sacc[2][r]);
            sacc[3][r] = __fmaf_rn(xv.w, u[r], sacc[3][r]);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < R; ++r) sacc[c][r] = __fadd_rn(sacc[c][r], __shfl_xor_sync(0xffffffffu, sacc[c][r], 16));
        if ((tid & 31) < 16) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < R; ++r) sm.red[(w * N + ln * 4 + c) * R + r] = sacc[c][r];
        }
        // U^T U partial of this warp: exact integers, one REDUX per entry
        int idx = 0;
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
          for (int r = j; r < R; ++r) {
            const int g = __reduce_add_sync(0xffffffffu, gacc[idx++]);
            if ((tid & 31) == 0) sm.gred[w * R * R + j * R + r] = g, sm.gred[w * R * R + r * R + j] = g;
          }
      }
      __syncthreads();
      for (int e = tid; e < N * R + R * R; e += NT) {  // fixed warp order, f64
        if (e < N * R) {
          double tot = 0.0;
#pragma unroll
          for (int w2 = 0; w2 < NW; ++w2) tot += (double)sm.red[w2 * N * R + e];
          sm.part[pbuf][e] = tot;
        } else {
          int g = 0;
#pragma unroll
          for (int w2 = 0; w2 < NW; ++w2) g += sm.gred[w2 * R * R + e - N * R];
          sm.part[pbuf][e] = (double)g;
        }
      }

      // ---------------- exchange partials across the cluster, every CTA sums in rank order ----------------
      if (cluster_size > 1) {
        cluster_barrier();
#ifndef LRFB_SIM
        for (int e = tid; e < N * R + R * R; e += NT) {
          double s = 0.0;
          for (int cr = 0; cr < cluster_size; ++cr) {
            const double* remote = cluster.map_shared_rank(&sm.part[pbuf][0], cr);
            s += remote[e];
          }
          if (e < N * R) sm.a2[e] = (float)s;
          else sm.b2[e - N * R] = (float)s;
        }
#endif
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
        gs_row_auto<R>(f, A, sm.b2, P.x_u8_range != 0, t2_native_v, P.lo, P.hi);
#pragma unroll
        for (int r = 0; r < R; ++r) sm.v[n * R + r] = f[r];
      }
      __syncthreads();
      // B = V^T V: V is integer-valued now, every order gives the same exact f32 result
      {
        const int w = tid >> 5, lane = tid & 31;
        for (int e = w; e < R * R; e += NW) {
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
    for (int i = tid; i < rows_here * R; i += NT) {
      const int row = i / R, r = i - row * R;
      const float val = bf16_to(sm.u[i]);
      if (P.U) P.U[(size_t)mat * M * R + (size_t)(row0 + row) * R + r] = val;
      if (P.Uq) P.Uq[(size_t)mat * P.uq_stride + (size_t)r * M + row0 + row] = (int8_t)(int)val;
    }
    if (crank == 0) {
      for (int i = tid; i < N * R; i += NT) {
        V[i] = sm.v[i];
        if (P.Vq) P.Vq[(size_t)mat * P.vq_stride + (size_t)(i % R) * N + i / R] = (int8_t)(int)sm.v[i];
      }
    }
  }
  if (cluster_size > 1) cluster_barrier();  // nobody leaves while its partials may still be read
}

}  // namespace lrfb

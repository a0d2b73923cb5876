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
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#endif

#ifndef LRFB_SIM

namespace lrfb {

constexpr int kTcRows = 768;       // largest row slice per CTA
constexpr int kTcMaxCluster = 16;  // 8 for the 768-row shape (portable), 16 for the 384-row shape (non-portable)
// Per-sweep exchange — no cluster barrier and no all-to-all of the partial sums on the critical path (DSMEM moves
// only ~20 B/cycle/SM, so a 2 KB x 8 all-to-all alone costs ~750 cycles):
//   1. reduce-scatter: rank c OWNS the 64/C columns n in [c*64/C, (c+1)*64/C).  After the tensor core has finished a
//      CTA's X^T U and U^T U partials, the four epilogue warps turn them into Q.24 fixed-point int64 (exact) and
//      PUSH (st.async) each column's sums to its owner only; the 4 x 4 U^T U partial goes to every rank.  Pushes
//      signal the receiver's `full` transaction mbarrier.
//   2. the owner adds the <= 8 slots (integers: order-free), runs the Gauss–Seidel update of its 64/C rows of V and
//      pushes each new row (16 B) to every rank (`vfull` mbarrier) — the all-gather.
//   3. two warps per CTA copy the 64 gathered rows into V and form V^T V (exact integers, REDUX).
// The receive buffers are single-buffered; a split cluster barrier (relaxed arrive after they were consumed, wait
// just before the next push — long satisfied by then) protects them.
template <int R, int ROWS, int NT>
struct TcSmem {
  static constexpr int N = 64;
  float x[ROWS * N];                    // swizzled f32 rows (A-phase)
  unsigned char ub[ROWS * 8];           // B operand: U as int8, K-major cores: (m/16)*128 + r*16 + m%16
  float v[N * R];
  float b[R * R];                       // V^T V of the float initialisation (first sweep of a matrix)
  int bi[2][16];                        // V^T V halves of the two gather warps (integer V, later sweeps)
  float s0inv[4];
  alignas(16) longlong2 srecv[128];     // [source rank][owned column][r pair]: Q.24 sums of X^T U
  alignas(16) int4 grecv[kTcMaxCluster * 4];  // [source rank][j]: row j of that rank's U^T U
  alignas(16) float4 vrecv[N];          // gathered new rows of V (padded to 4 columns)
  unsigned long long mma_done, clear_done, full, vfull, xfull[3];
  unsigned tmem_base;
  int next_mat;                         // rank 0: matrix index drawn for the whole cluster
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
__device__ __forceinline__ void tmem_ld4(unsigned taddr, unsigned (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ unsigned map_to_rank(unsigned cta_smem_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void push_2x64(unsigned remote, long long a, long long b, unsigned remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote),
               "l"(a), "l"(b), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void push_4x32(unsigned remote, int a, int b, int c, int d, unsigned remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
// relaxed: the slot reads this arrive publishes have already returned their values (they fed the V update that
// every thread passed a CTA barrier after), so no fence (MEMBAR.ALL.GPU + ERRBAR per thread and sweep) is needed
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// two independent single-rounded FMAs in one issue slot (FFMA2 with a broadcast scalar operand): bit-identical to
// fmaf(x, b.x, c.x) and fmaf(x, b.y, c.y)
__device__ __forceinline__ float2 ffma2_bcast(float x, float2 b, float2 c) {
  const float2 a = make_float2(x, x);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(rd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
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

// X rows in shared memory: two half planes [half][row][32 floats] in TMA's 128-byte swizzle (16-byte chunk c of a row
// sits at c ^ (row & 7)): what cp.async.bulk.tensor with CU_TENSOR_MAP_SWIZZLE_128B writes, and conflict-free both for
// the row-per-thread 16-byte reads of the U half-sweep and the column reads of the fixed-point conversion.
template <int ROWS>
__device__ __forceinline__ int x_off(int row, int ch /* 16-byte chunk 0..15 of the 64-float row */) {
  return (ch >> 3) * ROWS * 32 + row * 32 + (((ch & 7) ^ (row & 7)) << 2);
}

#ifdef LRFB_TC_TRACE
__device__ long long g_tc_trace[16 * 12];  // probe build only (tools/probes/tc_trace.cu): clock64 at 9 points of every sweep
#define TC_TRACE(pt) \
  if (blockIdx.x == 0 && tid == 0 && mats_seen == 1 && it < 12) g_tc_trace[(pt) * 12 + it] = clock64();
#define TC_TRACE_S(pt) \
  if (blockIdx.x == 0 && tid == 0 && mats_seen == 2) g_tc_trace[14 * 12 + (pt)] = clock64();
#else
#define TC_TRACE(pt)
#define TC_TRACE_S(pt)
#endif

template <int R, int ROWS, int NT, int MAXC = (ROWS > 384 ? 8 : 16)>
// MAXC: largest cluster the instantiation is launched with (sizes the unrolled slot loops).  768-row CTAs in clusters
// of 16 (non-portable size) keep matrices of up to 12 288 rows resident: kodim01's 10 292-row luma plane.
// 384-row shape: 2 CTAs per SM, so the exchange latency of one overlaps the arithmetic of the other.  Its 6 warps land
// 2,2,1,1 on the four scheduler partitions, so two CTAs put 4 warps on one partition's 16K registers: <= 128 registers
// per thread (bound declared as 256 threads x 2 CTAs).
__global__ void __launch_bounds__((ROWS <= 384 ? 256 : NT), (ROWS <= 384 ? 2 : 1))
bcd_tc_kernel(BcdBatch P, int cluster_size, int rows_per_cta, const __grid_constant__ CUtensorMap x_map, int use_tma) {
  constexpr int N = 64, RT = ROWS / NT, NW = NT / 32;
  constexpr int kTcRows = ROWS;
  constexpr int kTcColsA = ROWS / 4;                        // TMEM columns per A block (4 rows per cell)
  constexpr int kTcD1 = 2 * kTcColsA + 16, kTcD2 = kTcD1 + 16, kTcD3 = kTcD2 + 16;  // accumulator columns
  constexpr int kTmemCols = ROWS > 384 ? 512 : 256;
  constexpr int kMaxC = MAXC;  // largest cluster this instantiation is launched with
  static_assert(ROWS % NT == 0 && ROWS % 64 == 0 && NW >= 4 && kTcD3 + 8 <= kTmemCols, "shape");
  using S = TcSmem<R, ROWS, NT>;
  LRFB_DYN_SMEM(smem_raw);
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // same value, but provably warp-uniform (descriptor math in URs)
  const int M = P.M;
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int crank = (int)cluster.block_rank();
  const bool t2_native_u = bmm_native(R - 1, M, 1);
  constexpr bool t2_native_v = (long long)(R - 1) * N < 400;
  const bool from_a = P.s0 != nullptr;
  const int row0 = crank * rows_per_cta;
  const int rows_here = max(0, min(rows_per_cta, M - row0));
  unsigned mma_phase = 0, full_phase = 0, x_phase = 0;
  const int npr = N / cluster_size, npr_log = 6 - (31 - __clz(cluster_size));  // columns of V owned per rank
  int sweeps_done = 0;
#ifdef LRFB_TC_TRACE
  int mats_seen = 0;
#endif

  if (tid == 0) {
    mbar_init(&sm.mma_done, NW);  // one tcgen05.commit per warp and sweep
    mbar_init(&sm.clear_done, 1);
    mbar_init(&sm.full, 1);       // one arrive.expect_tx per sweep + the bytes of all ranks' pushes
    mbar_init(&sm.vfull, 1);
    for (int g = 0; g < 3; ++g) mbar_init(&sm.xfull[g], 1);  // TMA: expect_tx + the bytes of one row group
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
  const unsigned idesc_ss = idesc | (1u << 7);  // U^T U: A = the same int8 U tile read from shared memory (s8 x s8)
  // The accumulators are never cleared between sweeps: integer accumulation is exact and order-free, so every
  // warp issues the MMAs of its own rows as soon as they are projected, and the epilogue takes differences
  // (mod 2^32) against the previous sweep's totals.  One product against the all-zero B zeroes them here.
  int prev1[4], prev2[4], prev3[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) prev1[i] = prev2[i] = prev3[i] = 0;
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned long long bd0 = umma_desc(smem_u32(sm.ub), 128, 128);
    umma_i8_ts(tmem + kTcD1, tmem, bd0, idesc, 0);
    umma_i8_ts(tmem + kTcD2, tmem + kTcColsA, bd0, idesc, 0);
    umma_i8(tmem + kTcD3, umma_desc(smem_u32(sm.ub), 128, 0), bd0, idesc_ss, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.clear_done)) : "memory");
  }
  mbar_wait(&sm.clear_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster.sync();  // every peer's mbarriers are initialised before anyone pushes

  for (;;) {
    // ---- rank 0 draws the next matrix for the cluster ----
    if (sweeps_done > 0) cluster_wait();  // balance the last sweep's arrive: every exchange buffer is idle
    sweeps_done = 0;
    if (crank == 0 && tid == 0) sm.next_mat = atomicAdd(P.work_counter, 1);
    cluster.sync();
    const int mat = *cluster.map_shared_rank(&sm.next_mat, 0);
    if (mat >= P.n_mat) break;  // uniform over the cluster
    const float* X = P.X + (size_t)mat * P.x_stride;
    float* V = P.V + (size_t)mat * N * R;
#ifdef LRFB_TC_TRACE
    ++mats_seen;
#endif

    // ---- V first (the cluster barrier above separates it from the previous matrix), then this CTA's slice of X
    //      (swizzled f32, cp.async); the 64-step chains of V^T V run while the rows are in flight ----
    TC_TRACE_S(0)
    for (int i = tid; i < N * R; i += NT) sm.v[i] = V[i];
    if (tid < R) {
      float inv = 0.0f;
      if (from_a) {
        const float sv = P.s0[(size_t)mat * R + tid];
        inv = sv > 0.0f ? __fdiv_rn(1.0f, sv) : 0.0f;
      }
      sm.s0inv[tid] = inv;
    }
    __syncthreads();
    TC_TRACE_S(1)
    // three groups of ROWS/3 rows: the fixed-point conversion of a group starts as soon as it has landed.  TMA path:
    // one thread issues 2 tensor copies (32 columns x 256 rows, 32 KB) per group, rows past M are zero-filled by the
    // hardware (rows past this CTA's slice but inside M are loaded and ignored: their U rows are forced to zero).
    constexpr int kGroups = 3, kGroupRows = kTcRows / kGroups;
    static_assert(kTcRows % (kGroups * 64) == 0, "row groups are whole 64-row conversion chunks");
    if (use_tma) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of x vs the async writes
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const unsigned bar = smem_u32(&sm.xfull[g]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kGroupRows * 256) : "memory");
#pragma unroll
          for (int half = 0; half < 2; ++half)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                    "r"(smem_u32(&sm.x[half * kTcRows * 32 + g * kGroupRows * 32])),
                "l"(reinterpret_cast<unsigned long long>(&x_map)), "r"(half * 32), "r"(row0 + g * kGroupRows), "r"(mat), "r"(bar)
                : "memory");
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        for (int c = tid; c < kGroupRows * (N / 4); c += NT) {
          const int row = g * kGroupRows + (c >> 4), ch = c & 15;
          float* dst = &sm.x[x_off<ROWS>(row, ch)];
          if (row < rows_here) cp_async16(dst, X + (size_t)(row0 + row) * N + ch * 4);
          else dst[0] = dst[1] = dst[2] = dst[3] = 0.0f;
        }
        cp_async_commit();
      }
    }
    const float* Uinit = P.U + (size_t)mat * M * R + (size_t)row0 * R;
    TC_TRACE_S(2)
    gram_small<N, R>(sm.v, sm.b, tid);
    TC_TRACE_S(3)

    // ---- Q8.24 byte planes of X into tensor memory (A operand of the V-phase MMAs) ----
    {
      const int l = (warp & 3) * 32 + lane;     // accumulator / operand lane of this thread's TMEM quarter
      const int a_lo = l & 1, n = l >> 1;       // slice a_lo in block 0, a_lo + 2 in block 1 (lane pairs share n)
      // floor(x * 2^24) = hi16 << 16 | lo16 without the quarter-rate F2I: hi16 = floor(256 x) and lo16 = floor(65536 frac)
      // are read off the mantissas of round-toward-zero FMAs onto 2^23.  Slices 0/1 are the bytes 1/0 of hi16,
      // slices 2/3 those of lo16: block 0 (slice a_lo) comes from hi16, block 1 (slice a_lo + 2) from lo16.
      const unsigned sel = (1u - a_lo) | ((5u - a_lo) << 4);
      const unsigned lane_addr = tmem + ((unsigned)((warp & 3) * 32) << 16);
      const int sharers = (NW - (warp & 3) + 3) / 4;  // warps that own this TMEM lane quarter
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (use_tma) {
          mbar_wait(&sm.xfull[g], x_phase);
        } else {
          if (g == 0) cp_async_wait<2>();
          else if (g == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
          __syncthreads();
        }
        if (g == 0) { TC_TRACE_S(4) }
      for (int ch = warp >> 2; ch < kTcRows / 64; ch += sharers) {  // 64 rows (16 cells) per store
        if (ch / (kGroupRows / 64) != g) continue;
        unsigned w0[16], w1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          unsigned yh[4], yl[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int m = ch * 64 + c * 4 + t;
            const float xv = sm.x[x_off<ROWS>(m, n >> 2) + (n & 3)];
            const float h = __fmaf_rz(xv, 256.0f, 8388608.0f);                       // 2^23 + floor(256 x)
            const float frac = __fmaf_rn(xv, 256.0f, -__fadd_rn(h, -8388608.0f));     // exact, in [0, 1)
            yh[t] = __float_as_uint(h), yl[t] = __float_as_uint(__fmaf_rz(frac, 65536.0f, 8388608.0f));
          }
          w0[c] = __byte_perm(__byte_perm(yh[0], yh[1], sel), __byte_perm(yh[2], yh[3], sel), 0x5410);
          w1[c] = __byte_perm(__byte_perm(yl[0], yl[1], sel), __byte_perm(yl[2], yl[3], sel), 0x5410);
        }
        tmem_st16(lane_addr + ch * 16, w0);
        tmem_st16(lane_addr + kTcColsA + ch * 16, w1);
      }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      x_phase ^= 1;
    }
    TC_TRACE_S(5)
    // the first kReg A-phase rows of this thread stay in registers for all sweeps
    // the 2-CTAs-per-SM shape must stay within 128 registers; 256 threads x 3 rows have room for two register rows
    constexpr int kReg = (ROWS <= 384) ? 0 : (NT == 256 ? 2 : 1);
    constexpr int kRegN = kReg ? N : 1;
    float xr[kReg ? kReg : 1][kRegN];
#pragma unroll
    for (int i = 0; i < kReg; ++i) {
      const int row = tid + i * NT;
#pragma unroll
      for (int k4 = 0; k4 < N / 4; ++k4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&sm.x[x_off<ROWS>(row, k4)]);
        xr[i][(4 * k4 + 0) % kRegN] = t4.x, xr[i][(4 * k4 + 1) % kRegN] = t4.y;
        xr[i][(4 * k4 + 2) % kRegN] = t4.z, xr[i][(4 * k4 + 3) % kRegN] = t4.w;
      }
    }
    float uown[RT][R];  // this thread's U rows live in registers across the sweeps
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_TRACE_S(6)

    for (int it = 0; it < P.num_iters; ++it) {
      // ---------------- A-phase + Gauss–Seidel (identical arithmetic to bcd_resident_kernel) ----------------
      TC_TRACE(0)
      {
        float acc[RT][R];
        if (R % 2 == 0) {
          // packed chains: each FFMA2 advances two of the R ascending-k chains of a row
          constexpr int RP = R / 2 > 0 ? R / 2 : 1;
          float2 acc2[RT][RP];
#pragma unroll
          for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int r = 0; r < RP; ++r) acc2[i][r] = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int k4 = 0; k4 < N / 4; ++k4) {
            // the 4 x R values of V for these 4 k's are contiguous: R 16-byte loads (R = 2: two k's per load)
            float vflat[4 * R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(&sm.v[k4 * 4 * R + 4 * q]);
              vflat[4 * q] = t4.x, vflat[4 * q + 1] = t4.y, vflat[4 * q + 2] = t4.z, vflat[4 * q + 3] = t4.w;
            }
            float2 vk[4][RP];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int r = 0; r < RP; ++r) vk[k][r] = make_float2(vflat[k * R + 2 * r], vflat[k * R + 2 * r + 1]);
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
                xv = *reinterpret_cast<const float4*>(&sm.x[x_off<ROWS>(row, k4)]);
              }
#pragma unroll
              for (int r = 0; r < RP; ++r) {
                float2 a = acc2[i][r];
                a = ffma2_bcast(xv.x, vk[0][r], a);
                a = ffma2_bcast(xv.y, vk[1][r], a);
                a = ffma2_bcast(xv.z, vk[2][r], a);
                a = ffma2_bcast(xv.w, vk[3][r], a);
                acc2[i][r] = a;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int r = 0; r < RP; ++r) acc[i][(2 * r) % R] = acc2[i][r].x, acc[i][(2 * r + 1) % R] = acc2[i][r].y;
        } else {
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
              xv = *reinterpret_cast<const float4*>(&sm.x[x_off<ROWS>(row, k4)]);
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
        }
        TC_TRACE(1)
        float breg[R * R];  // V^T V: float chain of the initialisation in the first sweep, exact integers afterwards
#pragma unroll
        for (int e = 0; e < R * R; ++e) breg[e] = it == 0 ? sm.b[e] : (float)(sm.bi[0][e] + sm.bi[1][e]);
        float uden[R], urcp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) uden[r] = __fadd_rn(breg[r * R + r], kEps), urcp[r] = rcp_refined(uden[r]);
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
          // (one uniform branch per row instead of one per column: the flag is a compile-time constant inside)
          if (t2_native_u) gs_row_prepared<R>(f, acc[i], breg, uden, urcp, true, P.lo, P.hi);
          else gs_row_prepared<R>(f, acc[i], breg, uden, urcp, false, P.lo, P.hi);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (!ok) f[r] = 0.0f;
            uown[i][r] = f[r];
            // f is an integer in [-128, 127]: its two's-complement byte is the low mantissa byte of f + 1.5 * 2^23
            sm.ub[(row >> 4) * 128 + r * 16 + (row & 15)] = (unsigned char)__float_as_uint(__fadd_rn(f[r], 12582912.0f));
          }
          // ---- V-phase on the tensor core, issued per 32-row chunk as soon as the chunk is projected (the MMAs of
          //      chunk i run under the Gauss–Seidel arithmetic of chunk i+1): D_a += S_a^T U, D_3 += U^T U ----
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // int8 U rows -> visible to the tensor core
          __syncwarp();
          if (lane == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int j = warp_u + i * NW;  // k-chunk of 32 rows: rows i*NT + 32*warp ..
            const unsigned cbase = smem_u32(sm.ub) + j * 256;
            const unsigned long long bd = umma_desc(cbase, 128, 128);
            umma_i8_ts(tmem + kTcD1, tmem + j * 8, bd, idesc, 1);
            umma_i8_ts(tmem + kTcD2, tmem + kTcColsA + j * 8, bd, idesc, 1);
            // A = the same 8 x 32 int8 tile (stride 0 between the 8-row groups: accumulator lanes repeat mod 8)
            umma_i8(tmem + kTcD3, umma_desc(cbase, 128, 0), bd, idesc_ss, 1);
            if (i == RT - 1)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.mma_done)) : "memory");
          }
          __syncwarp();
        }
      }
      TC_TRACE(3)
      if (sweeps_done > 0) cluster_wait();  // every CTA has consumed the previous exchange (long satisfied)
      if (warp < 4) {
        // ---- epilogue: this sweep's sums = accumulator differences; Q.24 int64; reduce-scatter to the owners ----
        if (tid == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.full)), "r"(2048 + cluster_size * 64) : "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.vfull)), "r"(N * 16) : "memory");
        }
        mbar_wait(&sm.mma_done, mma_phase);
        TC_TRACE(4)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int l = warp * 32 + lane, a_lo = l & 1, n = l >> 1;
        unsigned d1[4], d2[4], d3[4];
        tmem_ld4(tmem + ((unsigned)(warp * 32) << 16) + kTcD1, d1);
        tmem_ld4(tmem + ((unsigned)(warp * 32) << 16) + kTcD2, d2);
        tmem_ld4(tmem + ((unsigned)(warp * 32) << 16) + kTcD3, d3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        TC_TRACE(10)
        long long tot[4];
        int e3[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e1 = (int)d1[r] - prev1[r], e2 = (int)d2[r] - prev2[r];  // exact mod 2^32
          e3[r] = (int)d3[r] - prev3[r];
          prev1[r] = (int)d1[r], prev2[r] = (int)d2[r], prev3[r] = (int)d3[r];
          const long long mine = ((long long)e1 << (24 - 8 * a_lo)) + ((long long)e2 << (8 - 8 * a_lo));
          tot[r] = mine + __shfl_xor_sync(0xffffffffu, mine, 1);  // all four slices of (n, r)
        }
        const long long p0 = a_lo ? tot[2] : tot[0], p1 = a_lo ? tot[3] : tot[1];
        TC_TRACE(11)
        const unsigned bar = smem_u32(&sm.full);
        const int owner = n >> npr_log, n_local = n & (npr - 1);
        push_2x64(map_to_rank(smem_u32(&sm.srecv[((crank << npr_log) + n_local) * 2 + a_lo]), owner), p0, p1,
                  map_to_rank(bar, owner));
        if (warp == 0) {  // U^T U row j = lane & 3 (accumulator lanes 0..3) to rank lane >> 2: one push per lane
          int row3[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) row3[r] = __shfl_sync(0xffffffffu, e3[r], lane & 3);
#pragma unroll
          for (int t = 0; t < kMaxC / 8; ++t) {
            const int cr = (lane >> 2) + 8 * t;
            if (cr < cluster_size)
              push_4x32(map_to_rank(smem_u32(&sm.grecv[crank * 4 + (lane & 3)]), cr), row3[0], row3[1], row3[2], row3[3],
                        map_to_rank(bar, cr));
          }
        }
      }
      mma_phase ^= 1;
      TC_TRACE(5)

      // ---------------- V update of the owned rows, then all-gather ----------------
      if (warp < 2 && warp * 32 < npr) {  // warps that own columns (warp 1 only when the cluster is a single CTA)
        mbar_wait(&sm.full, full_phase);
        TC_TRACE(6)
        // U^T U of the whole matrix: lane e sums entry e over the ranks, then every lane collects the R x R block
        // (unconditional loads from clamped slots, masked afterwards: all loads in flight at once, no branches)
        int ge = 0;
        {
          int gv[kMaxC];
#pragma unroll
          for (int cr = 0; cr < kMaxC; ++cr)
            gv[cr] = reinterpret_cast<const int*>(sm.grecv)[min(cr, cluster_size - 1) * 16 + (lane & 15)];
#pragma unroll
          for (int cr = 0; cr < kMaxC; ++cr) ge += cr < cluster_size ? gv[cr] : 0;
        }
        float b2[R * R];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
          for (int r = 0; r < R; ++r) b2[j * R + r] = (float)__shfl_sync(0xffffffffu, ge, j * 4 + r);
        float f[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (tid < npr) {
          long long s[4] = {0, 0, 0, 0};
          {
            longlong2 q0[kMaxC], q1[kMaxC];
#pragma unroll
            for (int cr = 0; cr < kMaxC; ++cr) {
              const int slot = ((min(cr, cluster_size - 1) << npr_log) + tid) * 2;
              q0[cr] = sm.srecv[slot], q1[cr] = sm.srecv[slot + 1];
            }
#pragma unroll
            for (int cr = 0; cr < kMaxC; ++cr) {
              const bool on = cr < cluster_size;
              s[0] += on ? q0[cr].x : 0, s[1] += on ? q0[cr].y : 0, s[2] += on ? q1[cr].x : 0, s[3] += on ? q1[cr].y : 0;
            }
          }
          TC_TRACE(12)
          const int n = crank * npr + tid;
          float fr[R], A[R];
#pragma unroll
          for (int r = 0; r < R; ++r) fr[r] = sm.v[n * R + r], A[r] = (float)((double)s[r] * 5.9604644775390625e-08);  // one rounding
          float vden[R], vrcp[R];
#pragma unroll
          for (int r = 0; r < R; ++r) vden[r] = __fadd_rn(b2[r * R + r], kEps), vrcp[r] = rcp_refined(vden[r]);
          gs_row_prepared<R>(fr, A, b2, vden, vrcp, t2_native_v, P.lo, P.hi);
#pragma unroll
          for (int r = 0; r < R; ++r) f[r] = fr[r];
        }
        // all-gather: (row k, rank c) pairs spread over the lanes — one or two pushes per lane instead of C per owner
        const unsigned vbar = smem_u32(&sm.vfull);
        if (npr > 32) {  // single-CTA cluster: every thread owns a row and keeps it local
          push_4x32(map_to_rank(smem_u32(&sm.vrecv[tid]), 0), __float_as_int(f[0]), __float_as_int(f[1]), __float_as_int(f[2]),
                    __float_as_int(f[3]), map_to_rank(vbar, 0));
        } else {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int q = lane + 32 * t, k = q & (npr - 1), cr = q >> npr_log;
            int w[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) w[r] = __shfl_sync(0xffffffffu, __float_as_int(f[r]), k);
            if (cr < cluster_size)
              push_4x32(map_to_rank(smem_u32(&sm.vrecv[crank * npr + k]), cr), w[0], w[1], w[2], w[3], map_to_rank(vbar, cr));
          }
        }
      }
      if (warp < 2) {
        mbar_wait(&sm.vfull, full_phase);
        TC_TRACE(13)
        const float4 row = sm.vrecv[tid];
        const float fv[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
        for (int r = 0; r < R; ++r) sm.v[tid * R + r] = fv[r];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
          for (int r = j; r < R; ++r) {  // V is integer-valued now: exact
            const int p = __reduce_add_sync(0xffffffffu, (int)fv[j] * (int)fv[r]);
            if (lane == 0) sm.bi[warp][j * R + r] = p, sm.bi[warp][r * R + j] = p;
          }
      }
      full_phase ^= 1;
      TC_TRACE(7)
      __syncthreads();
      TC_TRACE(8)
      cluster_arrive();  // this CTA is done with its receive buffers
      ++sweeps_done;
    }

    // ---- write the factors of this CTA's rows (and V once per cluster) ----
    TC_TRACE_S(7)
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
  // The loop exits right after a cluster.sync that followed the last exchange, so nobody is written to after exit; but
  // every CTA has just READ rank 0's matrix index through DSMEM, and rank 0 must not exit before those reads.
  cluster.sync();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace lrfb

#endif  // LRFB_SIM

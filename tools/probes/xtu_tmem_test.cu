// Dev probe: S = X^T U exactly on tcgen05 (kind::i8) with the A operand (byte slices of X) resident in TMEM
// and B = U (int8, K-major) in shared memory.  Validates the TMEM A layout (lane = MN row, 4 K-bytes per column).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "gram_i8.cuh"
using namespace lrfb;
constexpr int ROWS = 768;
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void umma_i8_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(160, 1) probe(const float* __restrict__ X, const signed char* __restrict__ U, double* __restrict__ S) {
  __shared__ __align__(128) unsigned char bsm[ROWS * 8];  // B: (k/16)*128 + n*16 + k%16
  __shared__ unsigned long long done;
  __shared__ unsigned tmem_base;
  __shared__ double comb[128 * 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base;
  if (warp < 4) {
    const int l = warp * 32 + lane, n = l & 63;
    for (int blk = 0; blk < 2; ++blk) {
      const int a = 2 * blk + (l >> 6);
      const int sh = 24 - 8 * a;
      for (int ch = 0; ch < ROWS / 64; ++ch) {  // 16 columns = 64 rows per store
        unsigned v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          unsigned w = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int m = ch * 64 + c * 4 + t;
            const unsigned fx = __float2uint_rz(X[m * 64 + n] * 16777216.0f);
            w |= ((fx >> sh) & 0xffu) << (8 * t);
          }
          v[c] = w;
        }
        tmem_st16(tmem + ((unsigned)(warp * 32) << 16) + blk * 192 + ch * 16, v);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int e = tid; e < ROWS * 8; e += 128) bsm[e] = 0;
  }
  __syncthreads();
  if (warp < 4)
    for (int m = tid; m < ROWS; m += 128)
      for (int r = 0; r < 4; ++r) bsm[(m >> 4) * 128 + r * 16 + (m & 15)] = (unsigned char)U[m * 4 + r];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4 && lane == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D = s32, A = u8 (TMEM), B = s8 K-major, N = 8, M = 128
    const unsigned idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned bbase = smem_u32(bsm);
    for (int j = 0; j < ROWS / 32; ++j) {
      const unsigned long long bd = umma_desc(bbase + j * 256, 128, 128);
      umma_i8_ts(tmem + 400, tmem + j * 8, bd, idesc, j > 0);
      umma_i8_ts(tmem + 416, tmem + 192 + j * 8, bd, idesc, j > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  if (warp < 4) {
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int l = warp * 32 + lane;
    unsigned d1[8], d2[8];
    tmem_ld8(tmem + ((unsigned)(warp * 32) << 16) + 400, d1);
    tmem_ld8(tmem + ((unsigned)(warp * 32) << 16) + 416, d2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int a = l >> 6;
    for (int r = 0; r < 4; ++r)
      comb[l * 4 + r] = (double)(int)d1[r] * exp2(-8.0 * a) + (double)(int)d2[r] * exp2(-8.0 * (a + 2));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 256) { const int n = tid >> 2, r = tid & 3; S[tid] = comb[n * 4 + r] + comb[(n + 64) * 4 + r]; }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  std::vector<float> hx(ROWS * 64); std::vector<signed char> hu(ROWS * 4);
  unsigned s = 777;
  for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) * (255.5f / 16777216.0f); }
  for (auto& v : hu) { s = s * 1664525u + 1013904223u; v = (signed char)((int)((s >> 10) % 32) - 16); }
  float* dx; signed char* du; double* ds;
  cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&du, hu.size()); cudaMalloc(&ds, 256 * 8);
  cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(du, hu.data(), hu.size(), cudaMemcpyHostToDevice);
  probe<<<1, 160>>>(dx, du, ds);
  cudaError_t err = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(err)); if (err) return 1;
  double hs[256]; cudaMemcpy(hs, ds, sizeof(hs), cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int n = 0; n < 64; ++n) for (int r = 0; r < 4; ++r) {
    long double acc = 0;
    for (int m = 0; m < ROWS; ++m) acc += (long double)(unsigned long long)(hx[m * 64 + n] * 16777216.0f) * hu[m * 4 + r];
    double ref = (double)(acc / 16777216.0L), got = hs[n * 4 + r];
    double rel = fabs(got - ref) / (fabs(ref) + 1e-30);
    if (rel > worst) worst = rel;
    if (r == 0) printf("%c", rel > 1e-12 ? 'x' : '.');
    if (rel > 1e-12 && r == 0 && (n == 32 || n == 16 || n == 8)) printf("\nn=%d r=%d got %.10g ref %.10g\n", n, r, got, ref);
  }
  printf("worst relative error %.3g\n", worst);
  return 0;
}

// TEST TOOLING — a minimal SIMT-on-CPU shim so the CUDA kernel sources in lrf_b200/csrc can be
// compiled with g++ (-DLRFB_SIM) and their index math / arithmetic checked against the oracle in
// the CPU-only container.  One OS thread per CUDA thread, std::barrier for __syncthreads, blocks
// run one after another.  It is never part of the product library (liblrfb.so is nvcc-built and
// has no CPU path); tests load the simulated library explicitly from tests/cpu_sim/.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_sim {
  unsigned x, y, z;
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

typedef int cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace sim {
struct Block {
  int nthreads;
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> scratch;
  std::vector<unsigned char> dyn;
  Block(int n, size_t smem) : nthreads(n), bar(n), scratch(n), dyn(smem + 64) {
    for (int w = 0; w * 32 < n; ++w) {
      int cnt = std::min(32, n - w * 32);
      warp_bar.emplace_back(new std::barrier<>(cnt));
    }
  }
};
struct Ctx {
  uint3_sim tid, bid;
  dim3 bdim, gdim;
  int lin;
  Block* blk;
};
inline thread_local Ctx ctx;

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
  int n = block.x * block.y * block.z;
  Block blk(n, smem);
  std::vector<std::thread> th;
  th.reserve(n);
  for (int t = 0; t < n; ++t) {
    th.emplace_back([&, t] {
      ctx.blk = &blk;
      ctx.bdim = block;
      ctx.gdim = grid;
      ctx.lin = t;
      ctx.tid = uint3_sim{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            ctx.bid = uint3_sim{bx, by, bz};
            kernel(args...);
            blk.bar.arrive_and_wait();
          }
    });
  }
  for (auto& x : th) x.join();
}

inline void sync_block() { ctx.blk->bar.arrive_and_wait(); }
inline void sync_warp() { ctx.blk->warp_bar[ctx.lin >> 5]->arrive_and_wait(); }
inline void* dyn_smem() {
  uintptr_t p = (uintptr_t)ctx.blk->dyn.data();
  return (void*)((p + 63) & ~(uintptr_t)63);
}

template <class T>
inline T exchange(T v, int src_lane) {  // every lane of the warp must call
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  Block& b = *ctx.blk;
  b.scratch[ctx.lin] = bits;
  sync_warp();
  int base = ctx.lin & ~31;
  int idx = base + src_lane;
  if (idx >= b.nthreads || src_lane < 0 || src_lane > 31) idx = ctx.lin;
  uint64_t got = b.scratch[idx];
  sync_warp();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
}  // namespace sim

#define threadIdx (sim::ctx.tid)
#define blockIdx (sim::ctx.bid)
#define blockDim (sim::ctx.bdim)
#define gridDim (sim::ctx.gdim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __syncthreads() sim::sync_block()
#define __syncwarp(...) sim::sync_warp()

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  int lane = sim::ctx.lin & 31;
  (void)width;
  return sim::exchange(v, lane ^ mask);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = sim::ctx.lin & 31;
  int base = lane & ~(width - 1);
  return sim::exchange(v, base + (src & (width - 1)));
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta, int width = 32) {
  int lane = sim::ctx.lin & 31;
  int src = lane + delta;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return sim::exchange(v, src);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    int got = __shfl_sync(0xffffffffu, (int)pred, l);
    int n = sim::ctx.blk->nthreads - (sim::ctx.lin & ~31);
    if (l < n && got) m |= 1u << l;
  }
  return m;
}
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31));
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta, int width = 32) {
  int lane = sim::ctx.lin & 31;
  int src = lane - delta;
  if (src < (lane & ~(width - 1))) src = lane;
  return sim::exchange(v, src);
}
inline int __reduce_max_sync(unsigned, int v) {
  for (int o = 16; o; o >>= 1) v = std::max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  for (int o = 16; o; o >>= 1) v = std::max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
inline int __reduce_min_sync(unsigned, int v) {
  for (int o = 16; o; o >>= 1) v = std::min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    unsigned got = __shfl_sync(0xffffffffu, v, l);
    int n = sim::ctx.blk->nthreads - (sim::ctx.lin & ~31);
    if (l < n && got == v) m |= 1u << l;
  }
  return m;
}
inline int atomicCAS(int* p, int cmp, int v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
inline int __reduce_add_sync(unsigned, int v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// IEEE single-operation intrinsics
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fdividef(float a, float b) { volatile float r = a / b; return r * (1.0f + 1.1920929e-07f); }  // perturbed by 1 ulp on purpose


inline float __frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int2float_rn(int a) { return (float)a; }
inline float __uint2float_rn(unsigned a) { return (float)a; }
inline int __float2int_rz(float a) { return (int)a; }
inline int __float2int_rn(float a) { return (int)std::nearbyintf(a); }
inline float __double2float_rn(double a) { return (float)a; }
inline float __ldg(const float* p) { return *p; }
inline double __ldg(const double* p) { return *p; }
inline unsigned char __ldg(const unsigned char* p) { return *p; }
inline signed char __ldg(const signed char* p) { return *p; }
inline int __ldg(const int* p) { return *p; }
using std::max;
using std::min;
using std::fmax;
using std::fmaxf;
using std::fmin;
using std::fminf;
using std::rintf;

// stream / memory plumbing used by the host side of the library
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "sim"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }

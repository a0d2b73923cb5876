// Dev probe: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}
template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float2 a[8];
  const float2 b = make_float2(1.0000001f, 0.9999999f), x = make_float2(threadIdx.x * 1e-3f, 1.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(i, -i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (PACKED) a[i] = ffma2(x, b, a[i]);
        else a[i].x = fmaf(x.x, b.x, a[i].x), a[i].y = fmaf(x.y, b.y, a[i].y);
      }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <bool PACKED>
void run(const char* name) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
  const int iters = 4096;
  k<PACKED><<<sms * 8, 256>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<PACKED><<<sms * 8, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = (double)sms * 8 * 256 * iters * 4 * 8 * 2 * 2;
  printf("%s: %.3f ms, %.1f TFLOP/s\n", name, ms, flops / ms / 1e9);
  cudaFree(out);
}
int main() {
  run<false>("FFMA  (scalar)");
  run<true>("FFMA2 (packed)");
  return 0;
}

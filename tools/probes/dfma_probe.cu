// Dev probe: vector FP64 (DFMA) and DMMA throughput on the GPU box.  nvcc -arch=sm_100a dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], 1.0000001, 1e-7);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345) out[threadIdx.x] = s;
}
__global__ void dmma(double* out, int iters) {
  double c[8][2]; for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[threadIdx.x] = s;
}
__global__ void f2f(double* out, const float* in, int iters) {
  double s = 0; float x = in[threadIdx.x & 31];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += (double)(x + (float)i); }
    x += 1.0f;
  }
  if (s == 1.2345) out[threadIdx.x] = s;
}
int main() {
  double* d; float* f; cudaMalloc(&d, 1 << 20); cudaMalloc(&f, 1024); cudaMemset(f, 0, 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148, iters = 20000; float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma<<<sms * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA: %.2f T fma/s (%.2f TFLOP/s)\n", sms * 4.0 * 256 * iters * 8 / ms / 1e9, 2 * sms * 4.0 * 256 * iters * 8 / ms / 1e9);
    cudaEventRecord(e0); dmma<<<sms * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA: %.2f TFLOP/s\n", 2.0 * sms * 4 * 8 * iters * 8 * 256 / ms / 1e9);
    cudaEventRecord(e0); f2f<<<sms * 4, 256>>>(d, f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("F2F.F64.F32 + DADD: %.2f T/s\n", sms * 4.0 * 256 * iters * 8 / ms / 1e9);
  }
  return 0;
}

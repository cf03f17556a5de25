// Micro-benchmark: FP64 DMMA (mma.sync m8n8k4) throughput per SM as a function of resident MMA warps
// per SM sub-partition and independent accumulator chains per warp.  One CTA per SM (forced by a
// large dynamic shared-memory request).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(double* out, int iters) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int warps, double* out, int sms) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<sms, warps * 32, 200 * 1024>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<CH><<<sms, warps * 32, 200 * 1024>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = (double)sms * warps * iters * CH * 512.;
  printf("warps/SM %2d (%d per sub-partition) chains %2d: %7.2f TFLOP/s\n", warps, warps / 4, CH, flops / ms * 1e-9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  for (int w : {4, 8, 12, 16}) { run<4>(w, out, sms); run<8>(w, out, sms); run<16>(w, out, sms); run<32>(w, out, sms); }
  return 0;
}

// Machine-peak probes used as roofline denominators by bench.py:
// FP64 tensor pipe (DMMA m8n8k4 register-resident loop), FP64 FMA pipe, HBM copy.
// MEASURED_PEAKS.json (driver-written) carries HBM and bf16 peaks only; the FP64 DMMA peak
// this path is bounded by has to be measured here.
#include "kernels.cuh"

namespace gxb {

namespace {

__global__ void __launch_bounds__(256) dmma_probe(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_probe(double* out, int iters) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_probe(const double4* __restrict__ in, double4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

template <typename F>
double time_ms(F&& f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms / reps;
}

}  // namespace

double probe_dmma_tflops(int iters) {
  const int blocks = 148 * 8;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  const double ms = time_ms([&] { dmma_probe<<<blocks, 256>>>(out, iters); }, 5);
  cudaFree(out);
  const double flops = double(blocks) * 8 /*warps*/ * iters * 8 * 512.;
  return flops / (ms * 1e-3) / 1e12;
}

double probe_dfma_tflops(int iters) {
  const int blocks = 148 * 8;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  const double ms = time_ms([&] { dfma_probe<<<blocks, 256>>>(out, iters); }, 5);
  cudaFree(out);
  const double flops = double(blocks) * 256 * iters * 8 * 2.;
  return flops / (ms * 1e-3) / 1e12;
}

double probe_copy_gbs(size_t bytes, int iters) {
  double4 *a, *b;
  bytes = bytes / 32 * 32;
  if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) return -1.;
  cudaMemset(a, 0, bytes);
  const size_t n = bytes / 32;
  const double ms = time_ms([&] { copy_probe<<<148 * 16, 512>>>(a, b, n); }, iters);
  cudaFree(a);
  cudaFree(b);
  return 2. * bytes / (ms * 1e-3) / 1e9;
}

}  // namespace gxb

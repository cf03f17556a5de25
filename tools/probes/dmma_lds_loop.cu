// Micro-benchmark: the fused kernel's MMA inner loop in isolation -- 64 x 16 warp tiles (8 x 2 DMMA
// m8n8k4 accumulators) fed by LDS.64 fragment loads from a 5-stage shared-memory ring laid out like
// FusedSmem (A[16][128] XOR-swizzled, P[16][68]); no barriers, no global traffic.  Reports the FP64
// tensor throughput of 8 (2 per sub-partition) and 4 MMA warps per SM, one CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int TP = 128, FK = 16, P_LD = 68, ST = 5;
struct Smem { double A[ST][FK][TP]; double P[ST][FK][P_LD]; };
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MI>
__device__ __forceinline__ void mma_stage(double (&acc)[8][2][2], const double* __restrict__ as,
                                          const double* __restrict__ ps, int a_ev, int a_od) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double a[MI], b[2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) a[mi] = as[kk * 4 * TP + ((mi & 1) ? a_od : a_ev) + (mi & ~1) * 8];
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) b[ni] = ps[kk * 4 * P_LD + ni * 8];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}
template <int MI>
__global__ void __launch_bounds__(256, 1) k(double* out, int nstages) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  for (int i = threadIdx.x; i < (int)(sizeof(Smem) / 8); i += blockDim.x) reinterpret_cast<double*>(raw)[i] = 1e-3 * (i & 7);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) & 1, wn = warp & 3;
  const int a_ev = (wm * 64 + g) ^ (t << 2), a_od = a_ev ^ 8;
  double acc[8][2][2];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;
  int s = 0;
  for (int it = 0; it < nstages; ++it) {
    mma_stage<MI>(acc, &S.A[s][t][0], &S.P[s][t][wn * 16 + g], a_ev, a_od);
    if (++s == ST) s = 0;
  }
  double r = 0;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) r += acc[mi][ni][0] + acc[mi][ni][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MI>
void run(int warps, double* out, int sms) {
  const int n = 20000;
  cudaFuncSetAttribute(k<MI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MI><<<sms, warps * 32, sizeof(Smem)>>>(out, n);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MI><<<sms, warps * 32, sizeof(Smem)>>>(out, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("MMA warps/SM %d, MI %d: %6.2f TFLOP/s  (%s)\n", warps, MI,
         (double)sms * warps * n * 4 * MI * 2 * 512. / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  run<8>(8, out, sms); run<8>(4, out, sms); run<6>(4, out, sms); run<4>(4, out, sms); run<2>(4, out, sms); run<4>(8, out, sms);
  return 0;
}

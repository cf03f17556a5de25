// Micro-benchmark: the fused kernel's MMA warps WITH the mbarrier ring around them -- 8 MMA warps
// (64 x 16 warp tiles from a 5-stage shared-memory ring) released / refilled by 4 producer warps that
// only do the handshake (no copies).  Variants isolate what the handshake costs the DMMA pipe:
//   MODE 0: no barriers at all (upper bound, = dmma_lds_loop)
//   MODE 1: full handshake, MMA arrives per warp (as in fused.cu)
//   MODE 2: full handshake, MMA warps consume TWO stages per wait/arrive (32 K rows)
//   MODE 3: as 1, but the producers are replaced by MMA warp 0 lane 0 re-arming the stage itself
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TP = 128, FK = 16, P_LD = 68, ST = 5;
struct Smem { double A[ST][FK][TP]; double P[ST][FK][P_LD]; uint64_t full[ST], empty[ST]; };
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    if (++spins > (1u << 22)) __trap();  // a protocol bug must not hang the GPU
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(su32(b)), "r"(ph) : "memory");
  }
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_stage(double (&acc)[8][2][2], const double* __restrict__ as,
                                          const double* __restrict__ ps, int a_ev, int a_od) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double a[8], b[2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) a[mi] = as[kk * 4 * TP + ((mi & 1) ? a_od : a_ev) + (mi & ~1) * 8];
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) b[ni] = ps[kk * 4 * P_LD + ni * 8];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(double* out, int nstages) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  for (int i = threadIdx.x; i < (int)(sizeof(S.A) + sizeof(S.P)) / 8; i += blockDim.x) reinterpret_cast<double*>(raw)[i] = 1e-3 * (i & 7);
  if (threadIdx.x == 0)
    for (int s = 0; s < ST; ++s) { mbar_init(&S.full[s], MODE == 3 ? 1 : 128); mbar_init(&S.empty[s], 8); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (warp >= 8) {
    if (MODE == 0 || MODE == 3) return;
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < nstages; ++it) {
      mbar_wait(&S.empty[s], ph ^ 1);
      mbar_arrive(&S.full[s]);
      if (++s == ST) { s = 0; ph ^= 1; }
    }
    return;
  }
  const int wm = warp >> 2, wn = warp & 3;
  const int a_ev = (wm * 64 + g) ^ (t << 2), a_od = a_ev ^ 8;
  double acc[8][2][2];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;
  int s = 0; uint32_t ph = 0;
  if (MODE == 3 && warp == 0 && lane == 0) for (int q = 0; q < ST; ++q) mbar_arrive(&S.full[q]);
  for (int it = 0; it < nstages; ++it) {
    if (MODE != 0) mbar_wait(&S.full[s], ph);
    mma_stage(acc, &S.A[s][t][0], &S.P[s][t][wn * 16 + g], a_ev, a_od);
    if (MODE == 2) {  // second half of a fat stage: one wait / arrive pair per 32 K rows
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[s]);
      if (++s == ST) { s = 0; ph ^= 1; }
      ++it;
      mma_stage(acc, &S.A[s][t][0], &S.P[s][t][wn * 16 + g], a_ev, a_od);  // filled together with its twin
    }
    if (MODE != 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[s]);
      if (MODE == 3) {
        // the last warp to release the stage would re-arm it in a real design; here warp 0 lane 0 waits
        // for the release and re-arms (costs that warp the wait)
        if (warp == 0 && lane == 0) { mbar_wait(&S.empty[s], ph); mbar_arrive(&S.full[s]); }
      }
    }
    if (++s == ST) { s = 0; ph ^= 1; }
  }
  double r = 0;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) r += acc[mi][ni][0] + acc[mi][ni][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(double* out, int sms) {
  const int n = 20000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, 384, sizeof(Smem)>>>(out, n);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms, 384, sizeof(Smem)>>>(out, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaDeviceSynchronize();
  printf("mode %d: %6.2f TFLOP/s  (%s)\n", MODE, (double)sms * 8 * n * 64 * 512. / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  setvbuf(stdout, nullptr, _IONBF, 0);
  run<0>(out, sms); run<1>(out, sms); run<3>(out, sms); run<2>(out, sms);
  return 0;
}

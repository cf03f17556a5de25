// Thin wrappers over the sm_100a PTX this path is built from: mbarrier producer/consumer
// pipelines, TMA tensor loads (cp.async.bulk.tensor), LDGSTS gathers and the FP64 tensor-core
// instruction (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gxb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on try_wait (which itself suspends the thread for a hardware time slice).  A protocol bug must not
// hang the GPU for ever, but a legitimate wait can be long under preemption / MPS / profiler replay, so the
// bound is generous: 2^26 failed probes (a minute or more), where a healthy wait takes a handful.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// all prior cp.async of this thread arrive on the barrier when they complete
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- LDGSTS ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* g, bool pred) {
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(g),
               "r"(sz)
               : "memory");
}

// ---- TMA ------------------------------------------------------------------------------
// 2-D tile load global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 2-D tile prefetch global -> L2 (no shared memory, no completion): hides the HBM latency of a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(tmap), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(tmap) : "memory");
}


// ---- 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256) ---------------------------------
// streamed once per stage of the pipeline: bypass L1 allocation
__device__ __forceinline__ void ldg256_stream(double (&v)[4], const double* p) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];\n"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p));
}
__device__ __forceinline__ void ldg128_stream(double (&v)[2], const double* p) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]),
               "d"(v[3])
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

// ---- L2 eviction-priority hints -------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void ldg256_stream_hint(double (&v)[4], const double* p, uint64_t pol) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;\n"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p), "l"(pol));
}
__device__ __forceinline__ void stg256_hint(double* p, const double (&v)[4], uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;\n" ::"l"(p), "d"(v[0]), "d"(v[1]),
               "d"(v[2]), "d"(v[3]), "l"(pol)
               : "memory");
}

// ---- DMMA -----------------------------------------------------------------------------
// D(8x8) += A(8x4,row) * B(4x8,col); lane = 4*g + t holds A[g][t], B[t][g], C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}


// ---- register re-partitioning between warp roles (all warps of a warpgroup must execute it) ---
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace gxb

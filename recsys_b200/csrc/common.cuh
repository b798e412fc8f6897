// Shared helpers for libctr_b200: error plumbing, device query, PTX wrappers
// (mbarrier, bulk async copy = TMA 1-D, vector reductions) for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../include/ctr_b200.h"

namespace ctr {

void set_error(const std::string& msg);
int fail_arg(const char* fn, const char* what);
int check_cuda(cudaError_t e, const char* fn);
int ensure_arch();      // CTR_OK or CTR_ERR_ARCH (cached per device)
int sm_count();         // multiprocessors of the current device (cached)
// A library-owned non-blocking stream + two timing-less events per device, for entry points that
// fork part of their work beside the caller's stream and join it back before returning (works
// under CUDA-graph capture: the fork / join become graph dependencies).  false on CUDA error.
bool aux_stream(cudaStream_t* stream, cudaEvent_t* fork, cudaEvent_t* join);
// Documented runtime options (ctr_set_option): value of `name`, or `dflt` when it was never set.
int option_get(const char* name, int dflt);

#define CTR_REQUIRE(cond, fn, what) \
  do {                              \
    if (!(cond)) return ::ctr::fail_arg(fn, what); \
  } while (0)
#define CTR_ARCH_OR_RETURN()          \
  do {                                \
    int _a = ::ctr::ensure_arch();    \
    if (_a != CTR_OK) return _a;      \
  } while (0)
#define CTR_LAUNCH_CHECK(fn) return ::ctr::check_cuda(cudaGetLastError(), fn)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Developer knobs (environment overrides of tcgen05 descriptor bits, tile schedules, launch modes)
// exist only in a build made with -DCTR_DEBUG_KNOBS=1 (CTR_DEBUG_KNOBS=1 python -m recsys_b200.build):
// the product library never reads them, so a stray variable cannot change - or corrupt - results.
#ifdef CTR_DEBUG_KNOBS
static inline const char* ctr_knob(const char* name) { return getenv(name); }
#else
static inline const char* ctr_knob(const char*) { return nullptr; }
#endif

// ------------------------------------------------------------------ device side
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 16-byte read-only load.
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// Streaming (evict-first) 16-byte load for data touched once.
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// Plain (coherent) 16-byte / 4-byte loads that the compiler may not sink below later code.
__device__ __forceinline__ float4 ld4_plain(const float* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld1_plain(const float* p) {
  float r;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// Vector float reduction into global memory: one 16-byte RED (sm_90+).
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// Device Adam schedule {t, lr_t, lr, block counter}: t = optimiser steps completed, kept as a
// uint32 BIT PATTERN in state[0] (a float counter would stop counting at 2^24 steps).
__device__ __forceinline__ unsigned adam_step_of(const float* state) { return __float_as_uint(state[0]); }
// System-scope flag words (peer-memory exchange): release store / acquire load, wall clock.
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait until *flag >= seq (flags are monotonic step numbers).  Bounded: after limit_ns it sets
// bit 0 of *err and returns, so a peer that never arrives cannot hang the GPU.
__device__ __forceinline__ void wait_flag_bounded(const int* flag, int seq, long long limit_ns, int* err) {
  if (ld_acquire_sys(flag) >= seq) return;
  const unsigned long long t0 = gtime_ns();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(64);
    if (gtime_ns() - t0 > static_cast<unsigned long long>(limit_ns)) {
      atomicOr(err, 1);
      return;
    }
  }
}
// lo part of the 3xTF32 split a.b ~ a_lo.b_hi + a_hi.b_lo + a_hi.b_hi.  tcgen05.mma.kind::tf32 reads
// the top 19 bits of an fp32 word, so the raw value serves as "hi"; lo = x - trunc_tf32(x) (exact),
// rounded to tf32 with integer ops (add half an ulp of the 10-bit mantissa; the tensor core drops
// the low 13 bits itself).
__device__ __forceinline__ float tcg_lo(float x) {
  const float l = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float(__float_as_uint(l) + 0x1000u);
}
__device__ __forceinline__ float4 tcg_lo4(float4 v) {
  return make_float4(tcg_lo(v.x), tcg_lo(v.y), tcg_lo(v.z), tcg_lo(v.w));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 b) {  // s*a + b
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4_shfl(float4 v, int src) {
  return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                     __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}

}  // namespace ctr

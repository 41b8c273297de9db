// common.cuh — shared host/device helpers for libstitchb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/stitch_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libstitchb200 is written for sm_100a (B200) only"
#endif

namespace sb {

// ----------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_device();          // SB_OK if the current device is sm_100
void count_launch(int n = 1);
unsigned int* debug_word_device();   // mapped host word for bounded-wait post-mortems (nullptr on failure)
int tune_get(int key, int dflt);   // sb_tune() override, else dflt

#define SB_REQUIRE(cond, code, ...)                 \
  do {                                              \
    if (!(cond)) {                                  \
      sb::set_error(__VA_ARGS__);                   \
      return (code);                                \
    }                                               \
  } while (0)

#define SB_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      sb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                    __FILE__, __LINE__);                                     \
      return SB_ECUDA;                                                       \
    }                                                                        \
  } while (0)

// Checks the launch that was just enqueued (no sync).
#define SB_LAUNCH_CHECK(name)                                                   \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      sb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));   \
      return SB_ECUDA;                                                          \
    }                                                                           \
    sb::count_launch();                                                         \
  } while (0)

#define SB_ENTER()                         \
  do {                                     \
    int _rc = sb::check_device();          \
    if (_rc != SB_OK) return _rc;          \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, and the library can be
// driven by several host threads of one process, one per GPU (the reference's nn.DataParallel,
// out.py:80 / evaluate.py:119).  One SmemOptIn per kernel family remembers, per device ordinal, the largest
// opt-in made so far; two threads racing on the same device both make the (idempotent) call.
class SmemOptIn {
 public:
  static constexpr int kMaxDevices = 64;
  // true when `bytes` exceeds what has been configured on the calling thread's current device
  bool need(size_t bytes, int* dev) {
    *dev = -1;
    if (cudaGetDevice(dev) != cudaSuccess || *dev < 0 || *dev >= kMaxDevices) return true;
    return set_[*dev].load(std::memory_order_acquire) < bytes;
  }
  void done(size_t bytes, int dev) {
    if (dev < 0 || dev >= kMaxDevices) return;
    size_t cur = set_[dev].load(std::memory_order_relaxed);
    while (cur < bytes && !set_[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {
    }
  }

 private:
  std::atomic<size_t> set_[kMaxDevices] = {};
};

inline cudaStream_t as_stream(sb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// --------------------------------------------------- exact fp32 arithmetic
// The library is compiled with -fmad=false; these wrappers make the intended
// rounding explicit where parity with the reference's separate ATen ops
// (one rounding per op) matters.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Streaming (read-once) global loads / stores.
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// grid_sample(align_corners=True) coordinate round trip, as the reference's
// Python + ATen's vectorised CPU kernel evaluate it (one rounding per op):
//   g  = 2*v / max(size-1,1) - 1        (core/warp_utils.py:74-75, core/utils/utils.py:66-67)
//   ix = (g + 1) * ((size-1)/2)         (ATen GridSamplerKernel.cpp ComputeLocation, align_corners)
// `den` = max(size-1,1) as float, `half` = (size-1)/2 as float.
__device__ __forceinline__ float grid_roundtrip(float v, float den, float half) {
  float g = fsub(fdiv(fmul(2.0f, v), den), 1.0f);
  return fmul(fadd(g, 1.0f), half);
}

// grid_roundtrip() for an INTEGER-valued `den` <= 2047 with the IEEE division restated as
// q = RN(a*r), q = fma(fma(-q, den, a), r, q), r = RN(1/den): bit-identical to div.rn for operands
// in the normal range (Markstein; proven exhaustively over all mantissas by
// tests/test_div_restatement.py), ~4 instructions instead of the ~15 of the generic division.
// RN(a / den) for an INTEGER-valued den <= 2047 with rcp = RN(1 / den): the Markstein restatement above.
__device__ __forceinline__ float div_small_int(float a, float den, float rcp) {
  float q = fmul(a, rcp);
  q = __fmaf_rn(__fmaf_rn(-q, den, a), rcp, q);
  const float aa = fabsf(a);
  if (!((aa > 1e-30f && aa < 1e30f) || aa == 0.0f)) q = fdiv(a, den);   // denormal range / huge / non-finite
  return q;
}
__device__ __forceinline__ float grid_roundtrip_rcp(float v, float den, float rcp, float half) {
  const float q = div_small_int(fmul(2.0f, v), den, rcp);
  return fmul(fadd(fsub(q, 1.0f), 1.0f), half);
}

}  // namespace sb

// Shared host/device helpers for libiiseg (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace iiseg {

// ---- error reporting -------------------------------------------------------
void set_error(const char* fmt, ...);
int32_t* diag_device_ptr();   // pinned, mapped host words the kernels write on a pipeline timeout
void count_launch(int n = 1);

#define IISEG_CHECK(cond, ...)                                 \
  do {                                                         \
    if (!(cond)) {                                             \
      ::iiseg::set_error(__VA_ARGS__);                         \
      return -1;                                               \
    }                                                          \
  } while (0)

#define IISEG_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      ::iiseg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                         cudaGetErrorString(e_));                                 \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

#define IISEG_LAUNCH_CHECK()                                                      \
  do {                                                                            \
    ::iiseg::count_launch();                                                      \
    cudaError_t e_ = cudaPeekAtLastError();                                       \
    if (e_ != cudaSuccess) {                                                      \
      ::iiseg::set_error("%s:%d launch -> %s", __FILE__, __LINE__,                \
                         cudaGetErrorString(e_));                                 \
      return -3;                                                                  \
    }                                                                             \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int kMaxDevices = 64;
int current_device();         // ordinal of the calling thread's device, clamped to [0, kMaxDevices)
int num_sms();                // SM count of that device
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: opt a kernel in once per device ordinal
#define IISEG_SMEM_OPT_IN(kernel, bytes)                                                            \
  do {                                                                                              \
    static bool configured_[::iiseg::kMaxDevices] = {false};                                        \
    const int dev_ = ::iiseg::current_device();                                                     \
    if (!configured_[dev_]) {                                                                       \
      IISEG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
      configured_[dev_] = true;                                                                     \
    }                                                                                               \
  } while (0)

// ---- device helpers --------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ---- packed bf16x2 helpers for the pool / ReLU epilogues --------------------------------------
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
// 0xFFFF in each 16-bit half whose values compare equal (IEEE ==: -0 == +0, NaN != NaN)
__device__ __forceinline__ uint32_t bf16x2_eq_mask(uint32_t a, uint32_t b) {
  return __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}
// Tie-mask word layout (8 channels x 4 window positions in one uint32): channel j = 2k + par
// (k = 32-bit word of the 16-byte vector, par = half), window position pos = 2*dy + dx:
//     bit = 16*par + 4*k + pos
// so the equality mask of one packed compare lands on its two bits with a single AND.
__device__ __forceinline__ uint32_t tie_bits(uint32_t eq_mask, int k, int pos) {
  return eq_mask & ((1u << (4 * k + pos)) | (1u << (16 + 4 * k + pos)));
}
// expands the two mask bits of (word k, position pos) back to 16-bit lane masks
__device__ __forceinline__ uint32_t tie_select(uint32_t bits, int k, int pos) {
  const uint32_t lo = (bits >> (4 * k + pos)) & 1u, hi = (bits >> (16 + 4 * k + pos)) & 1u;
  return (0u - lo) & 0x0000FFFFu | (0u - hi) & 0xFFFF0000u;
}

// exp of a non-positive argument for the channel softmax: ex2.approx(x * log2 e), ~2^-21 relative error.
// One definition shared by update.cu and the fused conv epilogue so both produce the same bits.
__device__ __forceinline__ float softmax_exp(float x) { return __expf(x); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace iiseg

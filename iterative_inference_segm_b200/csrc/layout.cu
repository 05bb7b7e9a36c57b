// Boundary layout conversion: the reference's tensors are NCHW float32
// (floatX, iterative_inference.py:29,234); the kernels work on NHWC bf16 with the
// channel count padded for 128-byte TMA rows.  These run once per batch, outside
// the iteration loop.  Thread order is pixel-fastest so the NCHW side is coalesced.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int C, int HW,
                                                   int C8, long long total, int split) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int pix = (int)(t % HW); t /= HW;
    const int cg = (int)(t % C8);
    const long long n = t / C8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      v[k] = c < C ? src[(n * C + c) * HW + pix] : 0.f;
    }
    const uint32_t hi[4] = {pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])};
    if (!split) {
      stg_v4(dst + (n * HW + pix) * C8 + cg, make_uint4(hi[0], hi[1], hi[2], hi[3]));
    } else {        // (hi | lo) bf16 pair of the fp32 value: hi = bf16(x), lo = bf16(x - hi)
      uint32_t lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lo[k] = pack_bf16x2(v[2 * k] - bf16_lo(hi[k]), v[2 * k + 1] - bf16_hi(hi[k]));
      stg_v4(dst + (n * HW + pix) * 2 * C8 + cg, make_uint4(hi[0], hi[1], hi[2], hi[3]));
      stg_v4(dst + (n * HW + pix) * 2 * C8 + C8 + cg, make_uint4(lo[0], lo[1], lo[2], lo[3]));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) unpack_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int HW,
                                                     int Cpad, long long total, int split) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int pix = (int)(t % HW); t /= HW;
    const int c = (int)(t % C);
    const long long n = t / C;
    if (!split) dst[i] = (float)src[(n * HW + pix) * Cpad + c];
    else dst[i] = (float)src[(n * HW + pix) * 2 * Cpad + c] + (float)src[(n * HW + pix) * 2 * Cpad + Cpad + c];
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? b : cap);
}

// NHWC bf16 (or the (hi | lo) pair) -> NHWC fp32, 8 channels per thread: the input of iiseg_channel_stats when the batch
// statistics of a rectified conv output are needed (DAE_h bn=1: DePool2D's mask pass, models/DAE_h.py).
__global__ void __launch_bounds__(256) widen_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, int C8, long long total,
                                                    int split) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C8;
    const int cg = (int)(i - pix * C8);
    const uint4 h = ldg_nc_v4(src + pix * (split ? 2 * C8 : C8) + cg);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = bf16_lo(hw[k]); v[2 * k + 1] = bf16_hi(hw[k]); }
    if (split) {
      const uint4 l = ldg_nc_v4(src + pix * 2 * C8 + C8 + cg);
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[2 * k] += bf16_lo(lw[k]); v[2 * k + 1] += bf16_hi(lw[k]); }
    }
    dst[i * 2] = make_float4(v[0], v[1], v[2], v[3]);
    dst[i * 2 + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

}  // namespace iiseg

extern "C" int iiseg_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int Cpad,
                                                int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(src && dst, "pack: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C && Cpad % 8 == 0, "pack: bad shape C=%d Cpad=%d", C, Cpad);
  const long long total = (long long)N * (Cpad / 8) * H * W;
  pack_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<uint4*>(dst), C, H * W, Cpad / 8, total, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_unpack_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int Cpad,
                                                  int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(src && dst, "unpack: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "unpack: bad shape");
  const long long total = (long long)N * C * H * W;
  unpack_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, C, H * W, Cpad, total, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_widen_nhwc_bf16_to_f32(const void* src, float* dst, long long pixels, int C, int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(src && dst, "widen: null tensor");
  IISEG_CHECK(pixels > 0 && C > 0 && C % 8 == 0, "widen: bad shape pixels=%lld C=%d", pixels, C);
  const long long total = pixels * (C / 8);
  widen_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<float4*>(dst), C / 8, total, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_unpack_nhwc_f32_to_nchw_f32(const float* src, float* dst, int N, int C, int H, int W, int Cpad,
                                                 void* stream) {
  using namespace iiseg;
  IISEG_CHECK(src && dst, "unpack: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "unpack: bad shape");
  const long long total = (long long)N * C * H * W;
  unpack_kernel<float><<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, C, H * W, Cpad,
                                                                                            total, 0);
  IISEG_LAUNCH_CHECK();
  return 0;
}

// Softmax tail + the iterative-inference update, fused into one streaming pass.
//
// Reference: the DAE's NHWC-reshape/softmax/NCHW tail (models/fcn_up.py:154-169)
// and the host loop body (iterative_inference.py:267-277):
//     grad = y - DAE(y, h);  y = clip(y - step*grad, 0, 1);
//     norm = mean_{pixels} ||grad||_2 over channels;  if norm < 1e-3: break
// One thread per pixel: 64 B of fp32 logits in (NHWC16), C coalesced NCHW plane
// reads/writes of the fp32 master y, one bf16 NHWC row out for the next conv.
// The norm is reduced deterministically: per-block partial sums in a fixed slot,
// summed in index order by iiseg_norm_finalize.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

constexpr int kUpdBlock = 256;
constexpr int kMaxC = 16;

__device__ __forceinline__ void load_logits16(const float* p, float* l) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = ldg_nc_v4(q + j);
    l[4 * j] = __uint_as_float(v.x); l[4 * j + 1] = __uint_as_float(v.y);
    l[4 * j + 2] = __uint_as_float(v.z); l[4 * j + 3] = __uint_as_float(v.w);
  }
}

// writes C probabilities (rest zero) as one bf16 NHWC row of Cpad channels
// split: the row is the (hi | lo) bf16 pair, 2*Cpad channels (fp32-accurate conv variant)
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* row, const float* v, int C, int Cpad, int split) {
  float t[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) t[c] = c < C ? v[c] : 0.f;
  uint32_t hi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) hi[k] = pack_bf16x2(t[2 * k], t[2 * k + 1]);
  uint4* o = reinterpret_cast<uint4*>(row);
  stg_v4(o, make_uint4(hi[0], hi[1], hi[2], hi[3]));
  stg_v4(o + 1, make_uint4(hi[4], hi[5], hi[6], hi[7]));
  for (int j = 2; j < Cpad / 8; ++j) stg_v4(o + j, make_uint4(0, 0, 0, 0));
  if (split) {
    uint32_t lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) lo[k] = pack_bf16x2(t[2 * k] - bf16_lo(hi[k]), t[2 * k + 1] - bf16_hi(hi[k]));
    o += Cpad / 8;
    stg_v4(o, make_uint4(lo[0], lo[1], lo[2], lo[3]));
    stg_v4(o + 1, make_uint4(lo[4], lo[5], lo[6], lo[7]));
    for (int j = 2; j < Cpad / 8; ++j) stg_v4(o + j, make_uint4(0, 0, 0, 0));
  }
}

__device__ __forceinline__ void softmax_c(const float* l, float* p, int C) {
  float m = l[0];
#pragma unroll
  for (int c = 1; c < kMaxC; ++c) if (c < C) m = fmaxf(m, l[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { p[c] = c < C ? softmax_exp(l[c] - m) : 0.f; s += p[c]; }
  const float inv = 1.0f / s;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) p[c] = __fmul_rn(p[c], inv);      // explicit roundings (no FMA contraction): the fused
                                                                      // conv epilogue (conv_epilogue16_update) must give the same bits
}

// grid = (nblk, N); kMode 0: softmax only (y = p), 1: softmax + update + norm, 2: grad = y - p (de_fn)
template <int kMode>
__global__ void __launch_bounds__(kUpdBlock) softmax_update_kernel(
    const float* __restrict__ logits, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
    float* __restrict__ p_out, const int32_t* __restrict__ active, float* __restrict__ norm_partial,
    int C, int HW, int Cpad, float step, const float* __restrict__ step_dev, int split) {
  const int n = blockIdx.y;
  if (kMode == 1 && step_dev != nullptr) step = __ldg(step_dev);
  if (kMode != 0 && active != nullptr && active[n] == 0) return;   // frozen image
  const int pix = blockIdx.x * kUpdBlock + threadIdx.x;
  float nrm = 0.f;
  if (pix < HW) {
    float l[kMaxC], p[kMaxC], yv[kMaxC];
    float* yb = y + (size_t)n * C * HW + pix;
    load_logits16(logits + ((size_t)n * HW + pix) * 16, l);
    if (kMode != 0) {            // the C plane reads of y go out together with the logits row: one memory latency, not two
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) yv[c] = c < C ? yb[(size_t)c * HW] : 0.f;
    }
    softmax_c(l, p, C);
    float out[kMaxC];
    if (kMode == 2) {        // de_fn: grad = y - DAE(y, h), y untouched
      float* gb = p_out + (size_t)n * C * HW + pix;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) if (c < C) gb[(size_t)c * HW] = yv[c] - p[c];
      return;
    } else if (kMode != 0) {
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        if (c < C) {
          const float yc = yv[c];
          const float g = __fsub_rn(yc, p[c]);
          ss = __fmaf_rn(g, g, ss);
          out[c] = fminf(fmaxf(__fsub_rn(yc, __fmul_rn(step, g)), 0.f), 1.f);
          yb[(size_t)c * HW] = out[c];
        } else out[c] = 0.f;
      }
      nrm = sqrtf(ss);
      if (p_out != nullptr) {
        float* pb = p_out + (size_t)n * C * HW + pix;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) if (c < C) pb[(size_t)c * HW] = p[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        out[c] = p[c];
        if (c < C) yb[(size_t)c * HW] = p[c];
      }
    }
    if (y_bf16 != nullptr) store_row_bf16(y_bf16 + ((size_t)n * HW + pix) * (split ? 2 * Cpad : Cpad), out, C, Cpad, split);
  }
  if (kMode == 2) return;
  if (kMode != 0) {
    __shared__ float red[kUpdBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < kUpdBlock / 32; ++w) s += red[w];
      norm_partial[(size_t)n * gridDim.x + blockIdx.x] = s;
    }
  }
}

__global__ void norm_finalize_kernel(const float* __restrict__ norm_partial, float* __restrict__ norm,
                                     int32_t* __restrict__ active, int32_t* __restrict__ n_exec, int nblk,
                                     float inv_hw, float eps, const float* __restrict__ eps_dev) {
  const int n = blockIdx.x;
  if (active[n] == 0) return;
  if (eps_dev != nullptr) eps = __ldg(eps_dev);
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += (double)norm_partial[(size_t)n * nblk + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float nv = (float)(tot * (double)inv_hw);
    norm[n] = nv;
    n_exec[n] += 1;
    if (nv < eps) active[n] = 0;
  }
}

// one thread per image: the fused conv epilogue already reduced ||g||_2 to one fixed-point word per image
__global__ void norm_finalize_fixed_kernel(unsigned long long* __restrict__ norm_acc, float* __restrict__ norm,
                                           int32_t* __restrict__ active, int32_t* __restrict__ n_exec, int N,
                                           double inv_hw, float eps, const float* __restrict__ eps_dev) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (eps_dev != nullptr) eps = __ldg(eps_dev);
  const unsigned long long acc = norm_acc[n];
  norm_acc[n] = 0ull;
  if (active[n] == 0) return;
  const float nv = (float)((double)acc * (1.0 / 1099511627776.0) * inv_hw);
  norm[n] = nv;
  n_exec[n] += 1;
  if (nv < eps) active[n] = 0;
}

}  // namespace iiseg

extern "C" int iiseg_update_blocks(int H, int W) { return (H * W + iiseg::kUpdBlock - 1) / iiseg::kUpdBlock; }

extern "C" int iiseg_softmax_nchw(const float* logits, float* p, void* y_bf16, int N, int C, int H, int W, int Cpad,
                                  int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(logits && p, "softmax: null tensor");
  IISEG_CHECK(N > 0 && C >= 1 && C <= kMaxC && H > 0 && W > 0, "softmax: bad shape");
  IISEG_CHECK(y_bf16 == nullptr || (Cpad >= 16 && Cpad % 8 == 0), "softmax: Cpad=%d", Cpad);
  dim3 grid(iiseg_update_blocks(H, W), N);
  softmax_update_kernel<0><<<grid, kUpdBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, p, reinterpret_cast<__nv_bfloat16*>(y_bf16), nullptr, nullptr, nullptr, C, H * W, Cpad, 0.f, nullptr, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_softmax_update(const float* logits, float* y, void* y_bf16, float* p_out, const int32_t* active,
                                    float* norm_partial, int N, int C, int H, int W, int Cpad, float step,
                                    const float* step_dev, int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(logits && y && norm_partial, "softmax_update: null tensor");
  IISEG_CHECK(N > 0 && C >= 1 && C <= kMaxC && H > 0 && W > 0, "softmax_update: bad shape");
  IISEG_CHECK(y_bf16 == nullptr || (Cpad >= 16 && Cpad % 8 == 0), "softmax_update: Cpad=%d", Cpad);
  dim3 grid(iiseg_update_blocks(H, W), N);
  softmax_update_kernel<1><<<grid, kUpdBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), p_out, active, norm_partial, C, H * W, Cpad, step, step_dev, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_softmax_grad(const float* logits, const float* y, float* grad, int N, int C, int H, int W,
                                  void* stream) {
  using namespace iiseg;
  IISEG_CHECK(logits && y && grad, "softmax_grad: null tensor");
  IISEG_CHECK(N > 0 && C >= 1 && C <= kMaxC && H > 0 && W > 0, "softmax_grad: bad shape");
  dim3 grid(iiseg_update_blocks(H, W), N);
  softmax_update_kernel<2><<<grid, kUpdBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, const_cast<float*>(y), nullptr, grad, nullptr, nullptr, C, H * W, 0, 0.f, nullptr, 0);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_norm_finalize(const float* norm_partial, float* norm, int32_t* active, int32_t* n_exec, int N,
                                   int H, int W, float eps, const float* eps_dev, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(norm_partial && norm && active && n_exec, "norm_finalize: null tensor");
  norm_finalize_kernel<<<N, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      norm_partial, norm, active, n_exec, iiseg_update_blocks(H, W), 1.0f / (float)(H * W), eps, eps_dev);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_norm_finalize_fixed(uint64_t* norm_acc, float* norm, int32_t* active, int32_t* n_exec, int N,
                                         int H, int W, float eps, const float* eps_dev, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(norm_acc && norm && active && n_exec, "norm_finalize_fixed: null tensor");
  norm_finalize_fixed_kernel<<<(N + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<unsigned long long*>(norm_acc), norm, active, n_exec, N, 1.0 / ((double)H * (double)W), eps, eps_dev);
  IISEG_LAUNCH_CHECK();
  return 0;
}

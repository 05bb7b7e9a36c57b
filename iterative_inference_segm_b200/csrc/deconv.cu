// Transposed convolution on <=16-channel score maps (FCN8 upsampling path).
//
// Reference: lasagne Deconv2DLayer(crop='valid', flip_filters=False, linear) at
// models/fcn8.py:90-91 (score2, k4 s2), :100-101 (score4, k4 s2), :109-110
// (upsample, k16 s8), plus the centre-cropped ElemwiseSumLayer that follows
// (:94-97, :104-107) and the final centre crop (:113-118).
//
//   out[oh,ow,co] = b[co] + sum_{ih,iw,ci} x[ih,iw,ci] * Wt[oh-ih*s][ow-iw*s][ci][co]
//
// These maps have 11 real channels: the work is (k/s)^2 * 16 * 16 FMAs per output
// pixel on a few hundred kB of data -- not GEMM-shaped, so it stays on the CUDA
// cores: one thread per output pixel, 16 fp32 accumulators, NHWC16 rows moved as
// 4 x 16-byte vectors, only the requested output window is computed; blocks are
// organised by output phase so the filter taps are staged in smem once per block.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

struct DeconvParams {
  const float* x; const float* w; const float* bias; const float* addend; float* out;
  int N, H, W, k, stride, oh0, ow0, OH, OW, AH, AW, ah0, aw0;
  int T;            // taps per axis that reach one output pixel: k / stride
  int ph_n, pw_n;   // output pixels per phase along h / w (upper bound)
};

// Output pixels whose full-resolution position has the same phase (fh % s, fw % s) use the same T x T
// filter taps (a = fh % s + t*s).  A block serves ONE phase: it stages the phase's T*T tap matrices
// (16 x 16 fp32 each) in shared memory once, and every thread then computes one output pixel from
// broadcast smem reads -- instead of each thread streaming its own 4 KB of taps through L1 (the 16x16
// upsampling filter bank is 256 KB: measured 1.7 ms for the three FCN8 deconvs, L2-bound).
// grid = (pixel chunks, s*s phases, N); block = 128 threads.
__global__ void __launch_bounds__(128) deconv16_kernel(const DeconvParams p) {
  extern __shared__ float4 s_w[];                        // [T*T][16 ci][4] float4 = [tap][ci][co]
  const int s = p.stride, T = p.T;
  const int py = blockIdx.y / s, px = blockIdx.y - py * s;
  for (int i = threadIdx.x; i < T * T * 64; i += 128) {
    const int tap = i >> 6, r = i & 63;
    const int a = py + (tap / T) * s, b = px + (tap % T) * s;
    s_w[i] = __ldg(reinterpret_cast<const float4*>(p.w + (static_cast<size_t>(a) * p.k + b) * 256) + r);
  }
  __syncthreads();
  // first window row / column with this phase, then every s-th
  const int oh_first = ((py - p.oh0) % s + s) % s, ow_first = ((px - p.ow0) % s + s) % s;
  const int idx = blockIdx.x * 128 + threadIdx.x;
  const int jh = idx / p.pw_n, jw = idx - jh * p.pw_n;
  const int oh = oh_first + jh * s, ow = ow_first + jw * s;
  if (oh >= p.OH || ow >= p.OW) return;
  const long long n = blockIdx.z;
  const int fh = oh + p.oh0, fw = ow + p.ow0;     // position in the full deconv output
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = __ldg(p.bias + c);
  for (int th = 0; th < T; ++th) {
    const int ih = fh / s - th;
    if (ih < 0 || ih >= p.H) continue;
    for (int tw = 0; tw < T; ++tw) {
      const int iw = fw / s - tw;
      if (iw < 0 || iw >= p.W) continue;
      const float4* xr = reinterpret_cast<const float4*>(p.x + ((n * p.H + ih) * p.W + iw) * 16);
      const float4* wr = s_w + (th * T + tw) * 64;
      float xv[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = __ldg(xr + j);
        xv[4 * j] = v.x; xv[4 * j + 1] = v.y; xv[4 * j + 2] = v.z; xv[4 * j + 3] = v.w;
      }
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wv = wr[ci * 4 + j];
          acc[4 * j] = fmaf(xv[ci], wv.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(xv[ci], wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xv[ci], wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(xv[ci], wv.w, acc[4 * j + 3]);
        }
      }
    }
  }
  if (p.addend != nullptr) {
    const float4* ar = reinterpret_cast<const float4*>(p.addend + ((n * p.AH + oh + p.ah0) * p.AW + ow + p.aw0) * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldg(ar + j);
      acc[4 * j] += v.x; acc[4 * j + 1] += v.y; acc[4 * j + 2] += v.z; acc[4 * j + 3] += v.w;
    }
  }
  float4* o = reinterpret_cast<float4*>(p.out + ((n * p.OH + oh) * p.OW + ow) * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
}

}  // namespace iiseg

extern "C" int iiseg_deconv2d_fwd(const iiseg_deconv_desc* d, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(d && d->x && d->weight && d->bias && d->out, "deconv: null tensor");
  IISEG_CHECK(d->k >= 1 && d->stride >= 1 && d->k >= d->stride, "deconv: bad filter k=%d stride=%d", d->k, d->stride);
  const int fullH = (d->H - 1) * d->stride + d->k, fullW = (d->W - 1) * d->stride + d->k;
  IISEG_CHECK(d->oh0 >= 0 && d->ow0 >= 0 && d->OH >= 1 && d->OW >= 1 && d->oh0 + d->OH <= fullH && d->ow0 + d->OW <= fullW,
              "deconv: window [%d+%d, %d+%d] outside %dx%d", d->oh0, d->OH, d->ow0, d->OW, fullH, fullW);
  if (d->addend != nullptr)
    IISEG_CHECK(d->ah0 >= 0 && d->aw0 >= 0 && d->ah0 + d->OH <= d->AH && d->aw0 + d->OW <= d->AW, "deconv: addend crop out of range");
  DeconvParams p;
  p.x = d->x; p.w = d->weight; p.bias = d->bias; p.addend = d->addend; p.out = d->out;
  p.H = d->H; p.W = d->W; p.k = d->k; p.stride = d->stride; p.oh0 = d->oh0; p.ow0 = d->ow0; p.OH = d->OH; p.OW = d->OW;
  p.AH = d->AH; p.AW = d->AW; p.ah0 = d->ah0; p.aw0 = d->aw0;
  IISEG_CHECK(d->k % d->stride == 0 && d->k / d->stride <= 4 && d->stride * d->stride <= 65535 && d->N <= 65535,
              "deconv: k=%d must be a multiple (<= 4x) of stride=%d", d->k, d->stride);
  p.N = d->N; p.T = d->k / d->stride;
  p.ph_n = (d->OH + d->stride - 1) / d->stride; p.pw_n = (d->OW + d->stride - 1) / d->stride;
  dim3 grid((p.ph_n * p.pw_n + 127) / 128, d->stride * d->stride, d->N);
  deconv16_kernel<<<grid, 128, p.T * p.T * 256 * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

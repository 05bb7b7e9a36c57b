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
// 4 x 16-byte vectors, only the requested output window is computed.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

struct DeconvParams {
  const float* x; const float* w; const float* bias; const float* addend; float* out;
  int H, W, k, stride, oh0, ow0, OH, OW, AH, AW, ah0, aw0;
  long long total;
};

__global__ void __launch_bounds__(128) deconv16_kernel(const DeconvParams p) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int ow = (int)(t % p.OW); t /= p.OW;
    const int oh = (int)(t % p.OH);
    const long long n = t / p.OH;
    const int fh = oh + p.oh0, fw = ow + p.ow0;   // position in the full deconv output
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = __ldg(p.bias + c);
    int ih_lo = (fh - p.k + p.stride) / p.stride; if (fh - p.k + 1 <= 0) ih_lo = 0;
    int ih_hi = fh / p.stride; if (ih_hi > p.H - 1) ih_hi = p.H - 1;
    int iw_lo = (fw - p.k + p.stride) / p.stride; if (fw - p.k + 1 <= 0) iw_lo = 0;
    int iw_hi = fw / p.stride; if (iw_hi > p.W - 1) iw_hi = p.W - 1;
    for (int ih = ih_lo; ih <= ih_hi; ++ih) {
      const int a = fh - ih * p.stride;
      for (int iw = iw_lo; iw <= iw_hi; ++iw) {
        const int b = fw - iw * p.stride;
        const float4* xr = reinterpret_cast<const float4*>(p.x + ((n * p.H + ih) * p.W + iw) * 16);
        const float4* wr = reinterpret_cast<const float4*>(p.w + ((size_t)(a * p.k + b)) * 256);
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
            const float4 wv = __ldg(wr + ci * 4 + j);
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
    float4* o = reinterpret_cast<float4*>(p.out + i * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
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
  p.total = (long long)d->N * d->OH * d->OW;
  long long blocks = (p.total + 127) / 128;
  const long long cap = (long long)num_sms() * 16;
  deconv16_kernel<<<(int)(blocks < cap ? blocks : cap), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

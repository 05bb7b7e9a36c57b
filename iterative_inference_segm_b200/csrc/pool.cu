// 2x2 max-pool with the tie-inclusive argmax mask, and the mask unpool (DePool2D).
//
// Reference: lasagne Pool2DLayer(2) (models/fcn_down.py:122) and DePool2D
// (layers/mylayers.py:88-115), whose "indices" are T.grad(pool, ones) = a 0/1
// mask of EVERY element equal to the window max.  Both kernels are pure
// streaming: NHWC bf16, one thread per (window, 8 channels), 16-byte vector
// accesses, channel-fastest thread order so a warp touches contiguous bytes.
//   pool  : reads 4 x 16 B, writes 16 B pooled + 4 B mask (8 nibbles)
//   unpool: one thread per (pooled pixel, 8 channels): reads 16 B + 4 B, writes 4 x 16 B
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

__global__ void __launch_bounds__(256) maxpool2_mask_kernel(const uint4* __restrict__ x, uint4* __restrict__ pooled,
                                                            uint32_t* __restrict__ mask, int H, int W, int C8,
                                                            int H2, int W2, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cg = (int)(t % C8); t /= C8;
    const int ow = (int)(t % W2); t /= W2;
    const int oh = (int)(t % H2);
    const long long n = t / H2;
    const long long row0 = ((n * H + 2 * oh) * W + 2 * ow) * C8 + cg;
    const long long row1 = row0 + (long long)W * C8;
    const uint4 v00 = ldg_nc_v4(x + row0), v01 = ldg_nc_v4(x + row0 + C8);
    const uint4 v10 = ldg_nc_v4(x + row1), v11 = ldg_nc_v4(x + row1 + C8);
    const uint32_t wa[4] = {v00.x, v00.y, v00.z, v00.w}, wb[4] = {v01.x, v01.y, v01.z, v01.w};
    const uint32_t wc[4] = {v10.x, v10.y, v10.z, v10.w}, wd[4] = {v11.x, v11.y, v11.z, v11.w};
    uint32_t mx[4], bits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mx[k] = bf16x2_max(bf16x2_max(wa[k], wb[k]), bf16x2_max(wc[k], wd[k]));
      bits |= tie_bits(bf16x2_eq_mask(wa[k], mx[k]), k, 0) | tie_bits(bf16x2_eq_mask(wb[k], mx[k]), k, 1) |
              tie_bits(bf16x2_eq_mask(wc[k], mx[k]), k, 2) | tie_bits(bf16x2_eq_mask(wd[k], mx[k]), k, 3);
    }
    stg_v4(pooled + i, make_uint4(mx[0], mx[1], mx[2], mx[3]));
    if (mask != nullptr) mask[i] = bits;
  }
}

// DePool2D restricted to an output window.  `u` is a dense
// [N,UH,UW,C] tensor whose element (0,0) sits at pooled-grid position (u_h0,u_w0); `out` is a
// dense [N,OH,OW,C] tensor whose element (0,0) is full-resolution pixel (o_h0,o_w0).  Output
// pixels in the trailing odd row / column of the HxW map (no pool window) are zero.
struct UnpoolParams {
  const uint4* u; const uint32_t* mask; uint4* out;
  int C8, UC8, MC8, H2, W2, UH, UW, u_h0, u_w0, OH, OW, o_h0, o_w0, PWN, PHN;   // MC8: channel groups of the mask (C8, or C8/2 for split pairs);
                                                                                // UC8: channel groups per pixel of u (C8, or 2*C8 when only the hi halves of a pair tensor are read)
};

// One thread per (pooled pixel touched by the window, 8 channels): u and the mask word are read ONCE
// and gate the (up to) 2x2 output pixels they feed.  grid = (ceil(PWN*C8/256), PHN, N): no 64-bit
// index arithmetic, consecutive threads walk the channels of a pixel, so every load and store
// instruction of a warp covers whole 128-byte lines.
__global__ void __launch_bounds__(256) unpool2_mask_kernel(const UnpoolParams p) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.PWN * p.C8) return;
  const int pc = idx / p.C8, cg = idx - pc * p.C8;
  const int ph = (p.o_h0 >> 1) + blockIdx.y, pw = (p.o_w0 >> 1) + pc;
  const size_t n = blockIdx.z;
  uint4 val = make_uint4(0, 0, 0, 0);
  uint32_t bits = 0;
  if (ph < p.H2 && pw < p.W2) {
    val = ldg_nc_v4(p.u + ((n * p.UH + (ph - p.u_h0)) * p.UW + (pw - p.u_w0)) * p.UC8 + cg);
    bits = __ldg(p.mask + ((n * p.H2 + ph) * p.W2 + pw) * p.MC8 + (cg >= p.MC8 ? cg - p.MC8 : cg));
  }
  const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const int oh = 2 * ph + dy - p.o_h0;
    if (oh < 0 || oh >= p.OH) continue;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int ow = 2 * pw + dx - p.o_w0;
      if (ow < 0 || ow >= p.OW) continue;
      const int pos = (dy << 1) | dx;
      uint32_t r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] = w[k] & tie_select(bits, k, pos);
      stg_v4(p.out + ((n * p.OH + oh) * p.OW + ow) * p.C8 + cg, make_uint4(r[0], r[1], r[2], r[3]));
    }
  }
}

static int stream_grid(long long total, int block) {
  long long blocks = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace iiseg

extern "C" int iiseg_maxpool2_mask_fwd(const void* x, void* pooled, uint32_t* mask, int N, int H, int W, int C,
                                       void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && pooled, "maxpool: null tensor");
  IISEG_CHECK(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, "maxpool: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int H2 = H / 2, W2 = W / 2, C8 = C / 8;
  const long long total = (long long)N * H2 * W2 * C8;
  maxpool2_mask_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(pooled), mask, H, W, C8, H2, W2, total);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_unpool2_mask_window_fwd(const void* u, const uint32_t* mask, void* out, int N, int H, int W, int C,
                                             int UH, int UW, int u_h0, int u_w0, int OH, int OW, int o_h0, int o_w0,
                                             int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(u && mask && out, "unpool: null tensor");
  IISEG_CHECK(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, "unpool: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int H2 = H / 2, W2 = W / 2;
  IISEG_CHECK(OH >= 1 && OW >= 1 && o_h0 >= 0 && o_w0 >= 0 && o_h0 + OH <= H && o_w0 + OW <= W,
              "unpool: output window [%d+%d, %d+%d] outside %dx%d", o_h0, OH, o_w0, OW, H, W);
  // every pool window the output touches must lie inside the u tensor
  const int ph_lo = o_h0 / 2, pw_lo = o_w0 / 2;
  int ph_hi = (o_h0 + OH - 1) / 2, pw_hi = (o_w0 + OW - 1) / 2;
  if (ph_hi > H2 - 1) ph_hi = H2 - 1;
  if (pw_hi > W2 - 1) pw_hi = W2 - 1;
  IISEG_CHECK(u_h0 <= ph_lo && u_w0 <= pw_lo && ph_hi < u_h0 + UH && pw_hi < u_w0 + UW,
              "unpool: u window [%d+%d, %d+%d] does not cover pooled rows %d..%d cols %d..%d", u_h0, UH, u_w0, UW,
              ph_lo, ph_hi, pw_lo, pw_hi);
  UnpoolParams p;
  p.u = reinterpret_cast<const uint4*>(u); p.mask = mask; p.out = reinterpret_cast<uint4*>(out);
  IISEG_CHECK(split >= 0 && split <= 2, "unpool: split must be 0, 1 or 2");
  p.MC8 = C / 8; p.C8 = split == 1 ? 2 * p.MC8 : p.MC8; p.UC8 = split ? 2 * p.MC8 : p.MC8; p.H2 = H2; p.W2 = W2; p.UH = UH; p.UW = UW; p.u_h0 = u_h0; p.u_w0 = u_w0;
  p.OH = OH; p.OW = OW; p.o_h0 = o_h0; p.o_w0 = o_w0;
  p.PWN = (o_w0 + OW - 1) / 2 - o_w0 / 2 + 1;     // pooled columns / rows the window touches (incl. a trailing odd one)
  p.PHN = (o_h0 + OH - 1) / 2 - o_h0 / 2 + 1;
  IISEG_CHECK(p.PHN <= 65535 && N <= 65535, "unpool: window too tall (%d pooled rows) or batch too large", p.PHN);
  dim3 grid(ceil_div(p.PWN * p.C8, 256), p.PHN, N);
  unpool2_mask_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_unpool2_mask_fwd(const void* u, const uint32_t* mask, void* out, int N, int H, int W, int C,
                                      void* stream) {
  return iiseg_unpool2_mask_window_fwd(u, mask, out, N, H, W, C, H / 2, W / 2, 0, 0, H, W, 0, 0, 0, stream);
}

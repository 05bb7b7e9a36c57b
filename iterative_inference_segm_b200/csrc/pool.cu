// 2x2 max-pool with the tie-inclusive argmax mask, and the mask unpool (DePool2D).
//
// Reference: lasagne Pool2DLayer(2) (models/fcn_down.py:122) and DePool2D
// (layers/mylayers.py:88-115), whose "indices" are T.grad(pool, ones) = a 0/1
// mask of EVERY element equal to the window max.  Both kernels are pure
// streaming: NHWC bf16, one thread per (window, 8 channels), 16-byte vector
// accesses, channel-fastest thread order so a warp touches contiguous bytes.
//   pool  : reads 4 x 16 B, writes 16 B pooled + 4 B mask (8 nibbles)
//   unpool: reads 16 B + 4 B, writes 4 x 16 B
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

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
    float a[8], b[8], c[8], d[8];
    unpack8(v00, a); unpack8(v01, b); unpack8(v10, c); unpack8(v11, d);
    float mx[8];
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mx[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
      const uint32_t nib = (a[k] == mx[k] ? 1u : 0u) | (b[k] == mx[k] ? 2u : 0u) |
                           (c[k] == mx[k] ? 4u : 0u) | (d[k] == mx[k] ? 8u : 0u);
      bits |= nib << (4 * k);
    }
    stg_v4(pooled + i, make_uint4(pack_bf16x2(mx[0], mx[1]), pack_bf16x2(mx[2], mx[3]),
                                  pack_bf16x2(mx[4], mx[5]), pack_bf16x2(mx[6], mx[7])));
    if (mask != nullptr) mask[i] = bits;
  }
}

// One thread per (ceil(H/2) x ceil(W/2) window, 8 channels): windows past the pooled extent
// only zero-fill the trailing odd row / column.
__global__ void __launch_bounds__(256) unpool2_mask_kernel(const uint4* __restrict__ u, const uint32_t* __restrict__ mask,
                                                           uint4* __restrict__ out, int H, int W, int C8, int H2,
                                                           int W2, int HC, int WC, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cg = (int)(t % C8); t /= C8;
    const int pw = (int)(t % WC); t /= WC;
    const int ph = (int)(t % HC);
    const long long n = t / HC;
    uint4 val = make_uint4(0, 0, 0, 0);
    uint32_t bits = 0;
    if (ph < H2 && pw < W2) {
      const long long src = ((n * H2 + ph) * W2 + pw) * C8 + cg;
      val = ldg_nc_v4(u + src);
      bits = __ldg(mask + src);
    }
    // expand each channel's nibble bit to a 16-bit lane mask
    const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int ih = 2 * ph + (pos >> 1), iw = 2 * pw + (pos & 1);
      if (ih >= H || iw >= W) continue;
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = (bits >> (4 * (2 * k) + pos)) & 1u;
        const uint32_t hi = (bits >> (4 * (2 * k + 1) + pos)) & 1u;
        const uint32_t sel = (lo ? 0x0000FFFFu : 0u) | (hi ? 0xFFFF0000u : 0u);
        o[k] = w[k] & sel;
      }
      stg_v4(out + ((n * H + ih) * W + iw) * C8 + cg, make_uint4(o[0], o[1], o[2], o[3]));
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

extern "C" int iiseg_unpool2_mask_fwd(const void* u, const uint32_t* mask, void* out, int N, int H, int W, int C,
                                      void* stream) {
  using namespace iiseg;
  IISEG_CHECK(u && mask && out, "unpool: null tensor");
  IISEG_CHECK(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, "unpool: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int H2 = H / 2, W2 = W / 2, HC = (H + 1) / 2, WC = (W + 1) / 2, C8 = C / 8;
  const long long total = (long long)N * HC * WC * C8;
  unpool2_mask_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(u), mask, reinterpret_cast<uint4*>(out), H, W, C8, H2, W2, HC, WC, total);
  IISEG_LAUNCH_CHECK();
  return 0;
}

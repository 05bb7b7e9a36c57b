// Streaming kernels around the convs of FC-DenseNet103 (models/FCDenseNet.py + FC_DenseNet.layers).
//
// The "Tiramisu" stack (ConcatLayer([stack, l]), models/FCDenseNet.py:84-89) is ONE fp32 NHWC tensor
// per dense block, allocated with its final channel count: a layer's 16 new feature maps are written
// behind the channels that exist so far, so no concat ever copies the stack.  Every BN_ReLU_Conv reads
// the first C channels through `bn_relu_pack` (BatchNormLayer with BATCH statistics,
// iterative_inference.py:187 batch_norm_use_averages=False, then rectify) into the zero-padded bf16
// NHWC tensor the tcgen05 conv kernel consumes.  Channel statistics are computed once per produced
// feature map (they do not depend on the consuming layer) by a deterministic two-level reduction.
// All kernels are bandwidth-bound: 16-byte accesses, channel-fastest thread order.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

// ---- BatchNorm (batch statistics) + rectify + bf16 pack ---------------------------------------------
// out[p, c] = bf16( relu( (x[p, c0+c] - mean[c]) * (gamma[c] * inv_std[c]) + beta[c] ) ), c < C; 0 for C <= c < Cpad.
// mean == NULL: plain copy/convert of the channel range (the deconv input of TransitionUp is not normalised).
// Thread t owns the 8-channel group t % C8pad for the whole launch (the launch makes the thread count a
// multiple of C8pad), keeps that group's BN coefficients in registers and walks pixels with 16-byte loads
// (two float4 in, one uint4 out per pixel).
__global__ void __launch_bounds__(256) bn_relu_pack_kernel(const float* __restrict__ x, long long P, int Cs, int c0, int C,
                                                           const float* __restrict__ mean, const float* __restrict__ inv_std,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int relu, uint4* __restrict__ out, int C8pad, int split) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const int cg = (int)(tid % C8pad);
  const long long pstep = nthr / C8pad;
  const int cb = cg * 8;
  const bool live = cb < C;                      // C is a multiple of 8 here: a group is all real or all padding
  float mu[8], sc[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = 0.f; sc[k] = 1.f; be[k] = 0.f;
    if (live && mean != nullptr) {
      mu[k] = __ldg(mean + cb + k); sc[k] = __fmul_rn(__ldg(gamma + cb + k), __ldg(inv_std + cb + k)); be[k] = __ldg(beta + cb + k);
    }
  }
  // four pixels per trip: eight independent 16-byte loads in flight per thread
  for (long long pix0 = tid / C8pad; pix0 < P; pix0 += 4 * pstep) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pix = pix0 + u * pstep;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f); b[u] = a[u];
      if (live && pix < P) {
        const float4* src = reinterpret_cast<const float4*>(x + pix * Cs + c0 + cb);
        a[u] = __ldg(src); b[u] = __ldg(src + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pix = pix0 + u * pstep;
      if (pix >= P) break;
      float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
      if (live) {
        if (mean != nullptr) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __fmaf_rn(__fsub_rn(v[k], mu[k]), sc[k], be[k]);
        }
        if (relu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
        }
      }
      const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]), h2 = pack_bf16x2(v[4], v[5]), h3 = pack_bf16x2(v[6], v[7]);
      if (!split) {
        stg_v4(out + pix * C8pad + cg, make_uint4(h0, h1, h2, h3));
      } else {        // (hi | lo) pair of the fp32 value: the operand format of the fp32-accurate convs (iiseg_conv_desc.split)
        stg_v4(out + pix * (2 * C8pad) + cg, make_uint4(h0, h1, h2, h3));
        stg_v4(out + pix * (2 * C8pad) + C8pad + cg,
               make_uint4(pack_bf16x2(v[0] - bf16_lo(h0), v[1] - bf16_hi(h0)), pack_bf16x2(v[2] - bf16_lo(h1), v[3] - bf16_hi(h1)),
                          pack_bf16x2(v[4] - bf16_lo(h2), v[5] - bf16_hi(h2)), pack_bf16x2(v[6] - bf16_lo(h3), v[7] - bf16_hi(h3))));
      }
    }
  }
}

// ---- per-channel batch statistics --------------------------------------------------------------------
// Level 1: a block sums x and x^2 over a chunk of kStatChunk pixels (fp32 per thread, 4 channels = one float4
// per load), combines its pixel lanes in fp64 in a fixed order and writes one (sum, sumsq) pair per (chunk, channel).
// Level 2: one thread per channel adds the chunks in index order (deterministic), biased variance,
// inv_std = 1/sqrt(var + eps) (lasagne BatchNormLayer, epsilon = 1e-4).
constexpr int kStatChunk = 2048;

// block = (bx, 256 / bx): x walks groups of 4 channels (one float4), y walks pixels of the chunk
__global__ void __launch_bounds__(256) channel_stats_partial_kernel(const float* __restrict__ x, long long P, int Cs, int c0, int C,
                                                                    double* __restrict__ partial) {
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index inside the channel range
  const int c = c4 * 4;
  const long long p0 = (long long)blockIdx.y * kStatChunk;
  const long long p1 = p0 + kStatChunk < P ? p0 + kStatChunk : P;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    // four loads in flight per thread; the summation order (pixel order per thread) is unchanged
    for (long long pb = p0 + threadIdx.y; pb < p1; pb += 4 * blockDim.y) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long p = pb + u * blockDim.y;
        v[u] = p < p1 ? __ldg(reinterpret_cast<const float4*>(x + p * Cs + c0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
        q[0] = __fmaf_rn(v[u].x, v[u].x, q[0]); q[1] = __fmaf_rn(v[u].y, v[u].y, q[1]);
        q[2] = __fmaf_rn(v[u].z, v[u].z, q[2]); q[3] = __fmaf_rn(v[u].w, v[u].w, q[3]);
      }
    }
  }
  __shared__ double sh[256][8];
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) { sh[t][k] = (double)s[k]; sh[t][4 + k] = (double)q[k]; }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    for (int yy = 0; yy < (int)blockDim.y; ++yy) {               // fixed order: deterministic
      const int tt = yy * blockDim.x + threadIdx.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ds[k] += sh[tt][k]; dq[k] += sh[tt][4 + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      partial[((size_t)blockIdx.y * C + c + k) * 2] = ds[k];
      partial[((size_t)blockIdx.y * C + c + k) * 2 + 1] = dq[k];
    }
  }
}

__global__ void __launch_bounds__(512) channel_stats_final_kernel(const double* __restrict__ partial, int n_chunks, int C, double inv_count, float eps,
                                                                  float* __restrict__ mean, float* __restrict__ inv_std) {
  // 32 channels per block, 16 warps: warp w adds chunks w, w+16, ... (coalesced double2 rows), then the 16 partial
  // sums are combined in warp order -- a fixed summation tree, so the statistics are run-to-run identical
  __shared__ double sh[16][32][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int k = w; k < n_chunks; k += 16) {
      const double2 v = *reinterpret_cast<const double2*>(partial + ((size_t)k * C + c) * 2);
      s += v.x; q += v.y;
    }
  }
  sh[w][lane][0] = s; sh[w][lane][1] = q;
  __syncthreads();
  if (w != 0 || c >= C) return;
  s = sh[0][lane][0]; q = sh[0][lane][1];
#pragma unroll
  for (int i = 1; i < 16; ++i) { s += sh[i][lane][0]; q += sh[i][lane][1]; }
  const double m = s * inv_count;
  double var = q * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  inv_std[c] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---- 2x2 max-pool on fp32 feature maps (TransitionDown), written into the next stack ------------------
__global__ void __launch_bounds__(256) maxpool2_f32_kernel(const float4* __restrict__ x, int H, int W, int C4s_in, int C4,
                                                           float4* __restrict__ out, int H2, int W2, int C4s_out, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cg = (int)(t % C4); t /= C4;
    const int ow = (int)(t % W2); t /= W2;
    const int oh = (int)(t % H2);
    const long long n = t / H2;
    const long long r0 = ((n * H + 2 * oh) * W + 2 * ow) * C4s_in + cg, r1 = r0 + (long long)W * C4s_in;
    const float4 a = x[r0], b = x[r0 + C4s_in], c = x[r1], d = x[r1 + C4s_in];
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)); m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)); m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
    out[((n * H2 + oh) * W2 + ow) * C4s_out + cg] = m;
  }
}

// ---- TransitionUp: interleave the four output-phase maps of the stride-2 3x3 transposed conv ----------
// out[n, oh, ow, c] = phase[(fh & 1) * 2 + (fw & 1)][n, fh >> 1, fw >> 1, c] with (fh, fw) = (oh + crop_h, ow + crop_w)
// the position in the (2H+1) x (2W+1) deconv output; phase maps are dense [N, H+1, W+1, Cp] fp32.
struct InterleaveParams { const float4* ph[4]; float4* out; int H1, W1, C4p, C4, OH, OW, crop_h, crop_w, C4s_out; long long total; };

__global__ void __launch_bounds__(256) deconv_interleave_kernel(const InterleaveParams p) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cg = (int)(t % p.C4); t /= p.C4;
    const int ow = (int)(t % p.OW); t /= p.OW;
    const int oh = (int)(t % p.OH);
    const long long n = t / p.OH;
    const int fh = oh + p.crop_h, fw = ow + p.crop_w;
    const float4* src = p.ph[((fh & 1) << 1) | (fw & 1)];
    p.out[((n * p.OH + oh) * p.OW + ow) * p.C4s_out + cg] = src[((n * p.H1 + (fh >> 1)) * p.W1 + (fw >> 1)) * p.C4p + cg];
  }
}

static int sgrid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? b : cap);
}

}  // namespace iiseg

extern "C" int iiseg_bn_relu_pack(const float* x, int N, int H, int W, int Cs, int c0, int C, const float* mean,
                                  const float* inv_std, const float* gamma, const float* beta, int relu, void* out,
                                  int Cpad, int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && out, "bn_relu_pack: null tensor");
  IISEG_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && c0 >= 0 && c0 + C <= Cs && Cpad >= C && Cpad % 8 == 0, "bn_relu_pack: bad shape C=%d Cs=%d Cpad=%d", C, Cs, Cpad);
  IISEG_CHECK(mean == nullptr || (inv_std && gamma && beta), "bn_relu_pack: incomplete BN parameters");
  IISEG_CHECK(C % 8 == 0 && c0 % 4 == 0 && Cs % 4 == 0, "bn_relu_pack: C must be a multiple of 8, c0 / Cs of 4");
  const long long P = (long long)N * H * W;
  const int C8 = Cpad / 8;
  int gcd = C8, r = 256;
  while (r) { const int t = gcd % r; gcd = r; r = t; }
  const int unit = C8 / gcd;                                  // grid must be a multiple of this: thread count % C8pad == 0
  long long blocks = (P * C8 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  blocks = (blocks + unit - 1) / unit * unit;
  bn_relu_pack_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, P, Cs, c0, C, mean, inv_std, gamma, beta, relu, reinterpret_cast<uint4*>(out), C8, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_channel_stats_chunks(int N, int H, int W) {
  const long long P = (long long)N * H * W;
  return (int)((P + iiseg::kStatChunk - 1) / iiseg::kStatChunk);
}

extern "C" int iiseg_channel_stats(const float* x, int N, int H, int W, int Cs, int c0, int C, float eps, double* scratch,
                                   float* mean, float* inv_std, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && scratch && mean && inv_std, "channel_stats: null tensor");
  IISEG_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && c0 >= 0 && c0 + C <= Cs, "channel_stats: bad shape");
  const long long P = (long long)N * H * W;
  const int n_chunks = iiseg_channel_stats_chunks(N, H, W);
  IISEG_CHECK(n_chunks <= 65535, "channel_stats: too many pixels");
  IISEG_CHECK(C % 4 == 0 && c0 % 4 == 0 && Cs % 4 == 0, "channel_stats: channel counts must be multiples of 4");
  const int bx = (C / 4) >= 16 ? 16 : 4;
  dim3 grid((C / 4 + bx - 1) / bx, n_chunks), block(bx, 256 / bx);
  channel_stats_partial_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, P, Cs, c0, C, scratch);
  IISEG_LAUNCH_CHECK();
  channel_stats_final_kernel<<<(C + 31) / 32, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(scratch, n_chunks, C, 1.0 / (double)P, eps,
                                                                                                  mean, inv_std);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_maxpool2_f32(const float* x, int N, int H, int W, int Cs_in, int C, float* out, int Cs_out, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && out, "maxpool2_f32: null tensor");
  IISEG_CHECK(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 4 == 0 && Cs_in % 4 == 0 && Cs_out % 4 == 0 && C <= Cs_in && C <= Cs_out, "maxpool2_f32: bad shape");
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)N * H2 * W2 * (C / 4);
  maxpool2_f32_kernel<<<sgrid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), H, W, Cs_in / 4, C / 4, reinterpret_cast<float4*>(out), H2, W2, Cs_out / 4, total);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_deconv_interleave(const float* p00, const float* p01, const float* p10, const float* p11, int N, int H,
                                       int W, int Cp, int C, int crop_h, int crop_w, float* out, int OH, int OW, int Cs_out,
                                       void* stream) {
  using namespace iiseg;
  IISEG_CHECK(p00 && p01 && p10 && p11 && out, "deconv_interleave: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && C % 4 == 0 && Cp % 4 == 0 && Cs_out % 4 == 0 && C <= Cp && C <= Cs_out, "deconv_interleave: bad channels");
  IISEG_CHECK(crop_h >= 0 && crop_w >= 0 && crop_h + OH <= 2 * H + 1 && crop_w + OW <= 2 * W + 1, "deconv_interleave: crop outside the %dx%d deconv output", 2 * H + 1, 2 * W + 1);
  InterleaveParams p;
  p.ph[0] = reinterpret_cast<const float4*>(p00); p.ph[1] = reinterpret_cast<const float4*>(p01);
  p.ph[2] = reinterpret_cast<const float4*>(p10); p.ph[3] = reinterpret_cast<const float4*>(p11);
  p.out = reinterpret_cast<float4*>(out);
  p.H1 = H + 1; p.W1 = W + 1; p.C4p = Cp / 4; p.C4 = C / 4; p.OH = OH; p.OW = OW; p.crop_h = crop_h; p.crop_w = crop_w; p.C4s_out = Cs_out / 4;
  p.total = (long long)N * OH * OW * p.C4;
  deconv_interleave_kernel<<<sgrid(p.total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

// Streaming kernels of the DAE training step (train_dae.py:243-335): everything between the tensor-core
// GEMMs of the backward pass.
//
//   forward  : noise + pack, then the inference kernels (full maps, no iteration hoists)
//   loss     : masked crossentropy + lmb * masked squared_error on softmax(logits) and its gradient
//              w.r.t. the logits (metrics.py:68-91,144-156)
//   backward : data gradients are the inference conv kernel on transposed / flipped filters (host side);
//              DePool2D backward = masked 2x2 sum; Pool2DLayer + rectify backward = the gradient to EVERY
//              tied maximum (Theano CPU MaxPoolGrad) where the pooled value is positive;
//              weight gradients are plain K-major GEMMs  dW[co][tap][ci] = sum_p g[p][co] * x[p+tap][ci]
//              on the same tensor-core kernel (a 1x1 "conv" whose channel axis is the pixel index), fed
//              by `transpose_shift_kernel`, which writes g^T and the nine tap-shifted x^T as bf16
//              [channels][pixels] matrices; a row of ones appended to x^T makes the bias gradient one
//              more output column.
//   update   : lasagne.updates.rmsprop on fp32 master weights kept in the GEMM layout, re-emitting the
//              bf16 forward filter bank and the flipped / transposed bank of the data-gradient conv.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

// ---- y + sigma * noise -> bf16 NHWC (GaussianNoiseLayer, models/fcn_down.py:60-67) -----------------
__global__ void __launch_bounds__(256) noise_pack_kernel(const float* __restrict__ y, const float* __restrict__ noise, float sigma,
                                                         uint4* __restrict__ dst, int C, int HW, int C8, long long total, int split) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int pix = (int)(t % HW); t /= HW;
    const int cg = (int)(t % C8);
    const long long n = t / C8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      v[k] = 0.f;
      if (c < C) {
        const long long j = (n * C + c) * HW + pix;
        v[k] = noise != nullptr ? __fmaf_rn(sigma, noise[j], y[j]) : y[j];
      }
    }
    const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]), h2 = pack_bf16x2(v[4], v[5]), h3 = pack_bf16x2(v[6], v[7]);
    if (!split) {
      stg_v4(dst + (n * HW + pix) * C8 + cg, make_uint4(h0, h1, h2, h3));
    } else {          // (hi | lo) pair of the fp32 value (the operand format of the fp32-accurate convs)
      uint4* row = dst + (n * HW + pix) * (2 * C8);
      stg_v4(row + cg, make_uint4(h0, h1, h2, h3));
      stg_v4(row + C8 + cg, make_uint4(pack_bf16x2(v[0] - bf16_lo(h0), v[1] - bf16_hi(h0)), pack_bf16x2(v[2] - bf16_lo(h1), v[3] - bf16_hi(h1)),
                                       pack_bf16x2(v[4] - bf16_lo(h2), v[5] - bf16_hi(h2)), pack_bf16x2(v[6] - bf16_lo(h3), v[7] - bf16_hi(h3))));
    }
  }
}

// ---- loss and d loss / d logits ------------------------------------------------------------------------
// sums[0] = sum_pix mask * CE, sums[1] = sum_pix mask, sums[2] = sum_pix m2 * mean_c (p-t)^2, sums[3] = sum_pix m2
// (fp64 atomics; the loss is sums[0]/sums[1] + lmb*sums[2]/sums[3]).  Pass 0 accumulates the sums, pass 1 reads
// the two denominators and writes dlogits (bf16 NHWC16):
//   dL/dp_k = -[k == true] * mask / (p_true * N_ce)   (0 where p_true was clipped)  +  lmb * 2 (p_k - t_k) m2 / (C * N_mse)
//   dL/dlogit_c = p_c * (dL/dp_c - sum_k dL/dp_k p_k)
// `terms` selects what train_dae.py:278-294 adds up: bit 0 crossentropy, bit 1 lmb * squared_error, bit 2 dice_loss
// (metrics.py:93-113) = -(2 I + 1) / (T + P + 1) on channel 1, I = sum t1 p1 (sums[4]), T = sum t1 (sums[5]), P = sum p1 (sums[6])
// over the entries whose target VALUE (int32 cast of the one-hot 0/1) differs from the void label id C:
//   dL/dp_1 += -(2 t1 (T + P + 1) - (2 I + 1)) / (T + P + 1)^2   for those entries.
constexpr int kTermCE = 1, kTermMSE = 2, kTermDice = 4;
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                        int C, int HW, float lmb, double* __restrict__ sums,
                                                        __nv_bfloat16* __restrict__ dlogits, int pass, int terms) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  double ce_s = 0.0, mk_s = 0.0, se_s = 0.0, m2_s = 0.0, di_s = 0.0, dt_s = 0.0, dp_s = 0.0;
  if (pix < HW) {
    float l[16], p[16], t[17];
    const uint4* q = reinterpret_cast<const uint4*>(logits + ((size_t)n * HW + pix) * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = ldg_nc_v4(q + j);
      l[4 * j] = __uint_as_float(v.x); l[4 * j + 1] = __uint_as_float(v.y); l[4 * j + 2] = __uint_as_float(v.z); l[4 * j + 3] = __uint_as_float(v.w);
    }
    float mx = l[0];
#pragma unroll
    for (int c = 1; c < 16; ++c) if (c < C) mx = fmaxf(mx, l[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { p[c] = c < C ? softmax_exp(l[c] - mx) : 0.f; s += p[c]; }
    const float inv = 1.0f / s;
    const float* tb = target + (size_t)n * (C + 1) * HW + pix;
    int tru = 0; float tbest = tb[0];
#pragma unroll
    for (int c = 0; c < 17; ++c) {
      if (c <= C) { t[c] = tb[(size_t)c * HW]; if (c > 0 && t[c] > tbest) { tbest = t[c]; tru = c; } }
    }
    const float mask = tru != C ? 1.f : 0.f;                  // void label = C (train_dae.py: void_labels of the iterator)
    const int idx = tru != C ? tru : 0;
    float m2 = 0.f, se = 0.f, ptrue = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < C) {
        p[c] *= inv;
        m2 += t[c];
        const float d = p[c] - t[c];
        se += d * d;
        if (c == idx) ptrue = p[c];
      }
    }
    const bool clipped = ptrue < 1e-7f || ptrue > 1.0f - 1e-7f;
    const float pc = fminf(fmaxf(ptrue, 1e-7f), 1.0f - 1e-7f);
    const int t1i = (int)t[1];                                  // T.cast(y_true_f, 'int32')
    const bool dice_keep = (terms & kTermDice) && t1i != C;     // T.neq(y_true_f, void_class[i]).nonzero()
    if (pass == 0) {
      ce_s = (double)(-logf(pc) * mask); mk_s = (double)mask; se_s = (double)(se / (float)C * m2); m2_s = (double)m2;
      if (dice_keep) { di_s = (double)((float)t1i * p[1]); dt_s = (double)t1i; dp_s = (double)p[1]; }
    } else {
      const float n_ce = (float)sums[1], n_mse = (float)sums[3];
      float dice_d = 0.f;
      if (dice_keep) {
        const double S = sums[5] + sums[6] + 1.0, I2 = 2.0 * sums[4] + 1.0;
        dice_d = (float)(-(2.0 * (double)t1i * S - I2) / (S * S));
      }
      float dp[16], dot = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        dp[c] = 0.f;
        if (c < C) {
          if (terms & kTermMSE) dp[c] = lmb * 2.f * (p[c] - t[c]) * m2 / ((float)C * n_mse);
          if ((terms & kTermCE) && c == idx && !clipped) dp[c] -= mask / (pc * n_ce);
          if (c == 1) dp[c] += dice_d;
          dot += dp[c] * p[c];
        }
      }
      float g[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) g[c] = c < C ? p[c] * (dp[c] - dot) : 0.f;
      uint4* o = reinterpret_cast<uint4*>(dlogits + ((size_t)n * HW + pix) * 16);
      stg_v4(o, make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7])));
      stg_v4(o + 1, make_uint4(pack_bf16x2(g[8], g[9]), pack_bf16x2(g[10], g[11]), pack_bf16x2(g[12], g[13]), pack_bf16x2(g[14], g[15])));
    }
  }
  if (pass == 0) {
    __shared__ double red[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ce_s += __shfl_xor_sync(0xffffffffu, ce_s, o); mk_s += __shfl_xor_sync(0xffffffffu, mk_s, o);
      se_s += __shfl_xor_sync(0xffffffffu, se_s, o); m2_s += __shfl_xor_sync(0xffffffffu, m2_s, o);
    }
    if ((threadIdx.x & 31) == 0) { const int w = threadIdx.x >> 5; red[0][w] = ce_s; red[1][w] = mk_s; red[2][w] = se_s; red[3][w] = m2_s; }
    __syncthreads();
    if (threadIdx.x < 4) {
      double a = 0.0;
      for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
      atomicAdd(sums + threadIdx.x, a);
    }
    if (terms & kTermDice) {          // block-uniform
      __syncthreads();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        di_s += __shfl_xor_sync(0xffffffffu, di_s, o); dt_s += __shfl_xor_sync(0xffffffffu, dt_s, o); dp_s += __shfl_xor_sync(0xffffffffu, dp_s, o);
      }
      if ((threadIdx.x & 31) == 0) { const int w = threadIdx.x >> 5; red[0][w] = di_s; red[1][w] = dt_s; red[2][w] = dp_s; }
      __syncthreads();
      if (threadIdx.x < 3) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
        atomicAdd(sums + 4 + threadIdx.x, a);
      }
    }
  }
}

// ---- DePool2D backward: g_u[ph,pw,c] = sum over the 2x2 window of mask * g_v ----------------------------
// g_v: dense window tensor [N,VH,VW,C] whose (0,0) is full-resolution pixel (v_h0,v_w0) (zero outside);
// g_u: dense [N,UH,UW,C] whose (0,0) is pooled position (u_h0,u_w0); mask: full [N,H2,W2,C/8].
struct DepoolBwdParams { const uint4* gv; const uint32_t* mask; uint4* gu; int C8, H2, W2, VH, VW, v_h0, v_w0, UH, UW, u_h0, u_w0; };

__global__ void __launch_bounds__(256) depool_bwd_kernel(const DepoolBwdParams p) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.UW * p.C8) return;
  const int uw = idx / p.C8, cg = idx - uw * p.C8;
  const int uh = blockIdx.y;
  const size_t n = blockIdx.z;
  const int ph = p.u_h0 + uh, pw = p.u_w0 + uw;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ph < p.H2 && pw < p.W2) {
    const uint32_t bits = __ldg(p.mask + ((n * p.H2 + ph) * p.W2 + pw) * p.C8 + cg);
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int vh = 2 * ph + (pos >> 1) - p.v_h0, vw = 2 * pw + (pos & 1) - p.v_w0;
      if (vh < 0 || vh >= p.VH || vw < 0 || vw >= p.VW) continue;
      const uint4 v = ldg_nc_v4(p.gv + ((n * p.VH + vh) * p.VW + vw) * p.C8 + cg);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t m = w[k] & tie_select(bits, k, pos);
        acc[2 * k] += bf16_lo(m); acc[2 * k + 1] += bf16_hi(m);
      }
    }
  }
  stg_v4(p.gu + ((n * p.UH + uh) * p.UW + uw) * p.C8 + cg, make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                                      pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7])));
}

// ---- Pool2DLayer(2) + rectify backward -------------------------------------------------------------------
// g_a[2ph+dy, 2pw+dx, c] = g_pool[ph,pw,c] if that element tied the window max (mask bit) and the max is > 0.
// rectify = 0.5*(x+|x|) has gradient 0.5 at exactly 0: in a window whose max is 0 (every element ties) the elements
// whose pre-rectifier value is exactly 0 (zmask bit) get g_pool/2, the negative ones 0 -- this is not a corner case:
// with zero-initialised biases the whole zero-padded border of every level is exactly 0.  Trailing odd row / column: 0.
__global__ void __launch_bounds__(256) pool_relu_bwd_kernel(const uint4* __restrict__ gpool, const uint4* __restrict__ pooled,
                                                            const uint32_t* __restrict__ mask, const uint32_t* __restrict__ zmask,
                                                            uint4* __restrict__ ga, int H, int W, int C8, int PW2) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= PW2 * C8) return;
  const int pw = idx / C8, cg = idx - pw * C8;            // pw in [0, ceil(W/2))
  const int ph = blockIdx.y;                              // in [0, ceil(H/2))
  const size_t n = blockIdx.z;
  const int H2 = H / 2, W2 = W / 2;
  uint32_t g[4] = {0, 0, 0, 0}, gh[4] = {0, 0, 0, 0}, bits = 0, zb = 0;
  if (ph < H2 && pw < W2) {
    const size_t pi = ((n * H2 + ph) * W2 + pw) * C8 + cg;
    const uint4 gv = ldg_nc_v4(gpool + pi), pv = ldg_nc_v4(pooled + pi);
    bits = __ldg(mask + pi);
    if (zmask != nullptr) zb = __ldg(zmask + pi);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, pwv[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lo = bf16_lo(pwv[k]) > 0.f ? 0x0000FFFFu : 0u, hi = bf16_hi(pwv[k]) > 0.f ? 0xFFFF0000u : 0u;
      g[k] = gw[k] & (lo | hi);                                                        // windows with a positive max
      gh[k] = pack_bf16x2(0.5f * bf16_lo(gw[k]), 0.5f * bf16_hi(gw[k])) & ~(lo | hi);  // windows whose max is 0: g/2 for exact zeros
    }
  }
#pragma unroll
  for (int pos = 0; pos < 4; ++pos) {
    const int oh = 2 * ph + (pos >> 1), ow = 2 * pw + (pos & 1);
    if (oh >= H || ow >= W) continue;
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = (g[k] & tie_select(bits, k, pos)) | (gh[k] & tie_select(zb, k, pos));
    stg_v4(ga + ((n * H + oh) * W + ow) * C8 + cg, make_uint4(r[0], r[1], r[2], r[3]));
  }
}

// ---- [pixels][channels] -> [channels][pixels] with a spatial shift (operands of the weight-gradient GEMM) ----
// out[(row0 + c) * ldo + p] = x[n, h0 + oh + dh, w0 + ow + dw, c0 + c]  (0 outside the HxW map), p = (n*OH + oh)*OW + ow,
// for c < C; columns p in [P, ldo) are written as zeros by the caller's memset.  64 pixels x 64 channels per block
// through shared memory so that both sides move 128-byte rows.
struct TransposeParams { const __nv_bfloat16* x; __nv_bfloat16* out; int N, H, W, Cs, c0, C, h0, w0, OH, OW, dh, dw, nshift; long long ldo, row0, shift_rows, P; };

__global__ void __launch_bounds__(256) transpose_shift_kernel(const TransposeParams p) {
  // 64 (+8) pixels x 64 channels per block.  33-word row pitch: the 4-byte stores of the load phase (lanes = 8 channel
  // chunks x 4 pixels) and the 2-byte column reads of the store phase (lanes = 4 pixel groups x 8 channels) are
  // bank-conflict free; global loads move whole 128-byte pixel rows, global stores 64-byte row segments.
  // nshift > 1: copy s of the output (rows + s * shift_rows) is the same matrix shifted by s pixels along the flat
  // pixel axis -- read once (8 extra pixels), written nshift times.
  __shared__ uint32_t tile[72][33];
  const long long p_base = (long long)blockIdx.x * 64;
  const int c_base = blockIdx.y * 64;
  const int n_pix = p.nshift > 1 ? 72 : 64;
  for (int idx = threadIdx.x; idx < n_pix * 8; idx += 256) {
    const int chunk = idx & 7, i = idx >> 3;                        // pixel i, channels chunk*8 .. +7
    const long long pp = p_base + i;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pp < p.P && c_base + chunk * 8 < p.C) {
      long long t = pp;
      const int ow = (int)(t % p.OW); t /= p.OW;
      const int oh = (int)(t % p.OH);
      const long long n = t / p.OH;
      const int ih = p.h0 + oh + p.dh, iw = p.w0 + ow + p.dw;
      if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = ldg_nc_v4(p.x + ((n * p.H + ih) * p.W + iw) * p.Cs + p.c0 + c_base + chunk * 8);
    }
    tile[i][chunk * 4 + 0] = v.x; tile[i][chunk * 4 + 1] = v.y; tile[i][chunk * 4 + 2] = v.z; tile[i][chunk * 4 + 3] = v.w;
  }
  __syncthreads();
  const unsigned short* t16 = reinterpret_cast<const unsigned short*>(&tile[0][0]);
  for (int sh = 0; sh < p.nshift; ++sh) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int pg = (idx & 3) + 4 * it, c = (idx >> 2) & 63;         // pixels pg*8 .. +7 of channel c
      if (c_base + c < p.C) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t lo = t16[(pg * 8 + 2 * j + sh) * 66 + c], hi = t16[(pg * 8 + 2 * j + 1 + sh) * 66 + c];
          w[j] = lo | (hi << 16);
        }
        stg_v4(p.out + (p.row0 + sh * p.shift_rows + c_base + c) * p.ldo + p_base + pg * 8, make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
  }
}

// ---- rmsprop + filter re-packing ---------------------------------------------------------------------------
// lasagne.updates.rmsprop (rho, eps): a <- rho*a + (1-rho)*g^2 ; w <- w - lr * g / sqrt(a + eps), on the fp32 master
// filter bank kept in the GEMM layout [Cout][taps][Cin_pad]; g is read from the weight-gradient GEMM output
// [Cout][ldg] (taps*Cin_pad filter columns, then the bias column at `bias_col`).  Emits the bf16 forward bank
// and, if wt != NULL, the bank of the data-gradient conv: wt[ci][R*S-1-tap][co] (flipped taps, transposed channels),
// for the ci range [ci0, ci0 + Ci_t) only (the concat conv propagates to its own half), padded to Co_pad columns.
// lasagne.updates.adam (the other optimiser train_dae.py:326-331 offers) runs in the same kernel when `m` is set:
// m <- beta1*m + (1-beta1)*g ; v <- beta2*v + (1-beta2)*g^2 (v lives in `acc`, beta2 in `rho`) ; w <- w - a_t * m / (sqrt(v) + eps),
// a_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) read from device memory (iiseg_adam_advance updates it once per step, so a captured
// graph replays with the right step count).
struct RmspropParams {
  float* w; float* acc; float* b; float* acc_b; const float* g; __nv_bfloat16* wb; __nv_bfloat16* wt;
  int Cout, taps, Cin_pad, ldg, bias_col, ci0, Ci_t, Co_pad, g_rstride; float lr, rho, eps; long long total;
  float* m; float* m_b; const float* a_t; float beta1;
};

__global__ void adam_advance_kernel(float* state, float lr, float beta1, float beta2) {
  const float t = state[0] + 1.f;                      // t_prev + 1 (float32, as lasagne's shared scalar)
  state[0] = t;
  state[1] = lr * sqrtf(1.f - powf(beta2, t)) / (1.f - powf(beta1, t));
}

__global__ void __launch_bounds__(256) rmsprop_pack_kernel(const RmspropParams p) {
  // One block = 64 output channels x 16 input channels of one tap.  Thread (r, q) owns 4 consecutive input channels
  // of output channel r: 16-byte accesses to w / acc / g, 8-byte stores of the forward bank.  The data-gradient
  // bank wants the OUTPUT channel contiguous, so the bf16 tile goes through shared memory and thread (ci, cq) writes
  // 4 consecutive output channels of input channel ci (8-byte stores, 128-byte runs per row).
  __shared__ unsigned short tile[64][18];
  const int n_cit = p.Cin_pad >> 4;
  const int tap = blockIdx.x / n_cit, ci_base = (blockIdx.x - tap * n_cit) << 4;
  const int co_base = blockIdx.y * 64;
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  const int co = co_base + r;
  uint32_t packed0 = 0, packed1 = 0;
  if (co < p.Cout) {
    const int rem = tap * p.Cin_pad + ci_base + 4 * q;
    const size_t i = ((size_t)co * p.taps) * p.Cin_pad + rem;
    const int fr = tap / 3;                                  // gradient column: filter row fr at stride g_rstride
    const float4 g = *reinterpret_cast<const float4*>(p.g + (size_t)co * p.ldg + fr * p.g_rstride + (rem - fr * 3 * p.Cin_pad));
    float4 a = *reinterpret_cast<const float4*>(p.acc + i);
    float4 w = *reinterpret_cast<const float4*>(p.w + i);
    if (p.m != nullptr) {          // adam
      const float at = __ldg(p.a_t), b1 = p.beta1;
      float4 m = *reinterpret_cast<const float4*>(p.m + i);
      m.x = b1 * m.x + (1.f - b1) * g.x; a.x = p.rho * a.x + (1.f - p.rho) * g.x * g.x; w.x -= at * m.x / (sqrtf(a.x) + p.eps);
      m.y = b1 * m.y + (1.f - b1) * g.y; a.y = p.rho * a.y + (1.f - p.rho) * g.y * g.y; w.y -= at * m.y / (sqrtf(a.y) + p.eps);
      m.z = b1 * m.z + (1.f - b1) * g.z; a.z = p.rho * a.z + (1.f - p.rho) * g.z * g.z; w.z -= at * m.z / (sqrtf(a.z) + p.eps);
      m.w = b1 * m.w + (1.f - b1) * g.w; a.w = p.rho * a.w + (1.f - p.rho) * g.w * g.w; w.w -= at * m.w / (sqrtf(a.w) + p.eps);
      *reinterpret_cast<float4*>(p.m + i) = m;
    } else {
      a.x = p.rho * a.x + (1.f - p.rho) * g.x * g.x; w.x -= p.lr * g.x / sqrtf(a.x + p.eps);
      a.y = p.rho * a.y + (1.f - p.rho) * g.y * g.y; w.y -= p.lr * g.y / sqrtf(a.y + p.eps);
      a.z = p.rho * a.z + (1.f - p.rho) * g.z * g.z; w.z -= p.lr * g.z / sqrtf(a.z + p.eps);
      a.w = p.rho * a.w + (1.f - p.rho) * g.w * g.w; w.w -= p.lr * g.w / sqrtf(a.w + p.eps);
    }
    *reinterpret_cast<float4*>(p.acc + i) = a;
    *reinterpret_cast<float4*>(p.w + i) = w;
    packed0 = pack_bf16x2(w.x, w.y); packed1 = pack_bf16x2(w.z, w.w);
    *reinterpret_cast<uint2*>(p.wb + i) = make_uint2(packed0, packed1);
    if (rem == 0) {       // one thread per output channel also updates the bias
      const float gb = p.g[(size_t)co * p.ldg + p.bias_col];
      const float ab = p.rho * p.acc_b[co] + (1.f - p.rho) * gb * gb;
      p.acc_b[co] = ab;
      if (p.m != nullptr) {
        const float mb = p.beta1 * p.m_b[co] + (1.f - p.beta1) * gb;
        p.m_b[co] = mb;
        p.b[co] -= __ldg(p.a_t) * mb / (sqrtf(ab) + p.eps);
      } else {
        p.b[co] -= p.lr * gb / sqrtf(ab + p.eps);
      }
    }
  }
  if (p.wt == nullptr || ci_base < p.ci0 || ci_base >= p.ci0 + p.Ci_t) return;      // block-uniform
  tile[r][4 * q + 0] = (unsigned short)(packed0 & 0xFFFFu); tile[r][4 * q + 1] = (unsigned short)(packed0 >> 16);
  tile[r][4 * q + 2] = (unsigned short)(packed1 & 0xFFFFu); tile[r][4 * q + 3] = (unsigned short)(packed1 >> 16);
  __syncthreads();
  const int ci = threadIdx.x >> 4, cq = threadIdx.x & 15;
  const int co4 = co_base + 4 * cq;
  if (co4 < p.Cout) {       // Cout % 4 == 0 (padded channel counts)
    const uint32_t lo = (uint32_t)tile[4 * cq][ci] | ((uint32_t)tile[4 * cq + 1][ci] << 16);
    const uint32_t hi = (uint32_t)tile[4 * cq + 2][ci] | ((uint32_t)tile[4 * cq + 3][ci] << 16);
    *reinterpret_cast<uint2*>(p.wt + ((size_t)(ci_base + ci - p.ci0) * p.taps + (p.taps - 1 - tap)) * p.Co_pad + co4) = make_uint2(lo, hi);
  }
}

// out[i] = sum_s in[s][i]  (the K slabs of a split-K weight-gradient GEMM, summed in slab order: deterministic)
__global__ void __launch_bounds__(256) sum_slabs_kernel(const float4* __restrict__ in, float4* __restrict__ out, int S, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = in[i];
    for (int s = 1; s < S; ++s) {
      const float4 b = in[(long long)s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// ---- bias gradient: channel sums of a bf16 NHWC gradient, two deterministic stages ---------------------------------
__global__ void __launch_bounds__(256) bias_grad_partial_kernel(const uint4* __restrict__ g, long long P, int C8, float* __restrict__ part) {
  // block = 256 threads = (256 / C8l) pixel lanes x C8l channel groups, C8l = min(C8, 32) per pass
  const int chunk = blockIdx.x, chunks = gridDim.x;
  const long long per = (P + chunks - 1) / chunks;
  const long long p0 = chunk * per, p1 = (p0 + per < P) ? p0 + per : P;
  __shared__ float red[256][9];
  for (int cg0 = 0; cg0 < C8; cg0 += 32) {
    const int ncg = (C8 - cg0) < 32 ? (C8 - cg0) : 32;
    const int lanes = 256 / ncg;
    const int cg = threadIdx.x % ncg, pl = threadIdx.x / ncg;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (pl < lanes) {
      // four independent 16-byte loads in flight per thread; the additions keep the order q, q + lanes, q + 2 lanes, ...
      const uint4* gp = g + cg0 + cg;
      long long q = p0 + pl;
      for (; q + 3LL * lanes < p1; q += 4LL * lanes) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_nc_v4(gp + (q + (long long)u * lanes) * C8);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { a[2 * j] += bf16_lo(w[j]); a[2 * j + 1] += bf16_hi(w[j]); }
        }
      }
      for (; q < p1; q += lanes) {
        const uint4 v = ldg_nc_v4(gp + q * C8);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[2 * j] += bf16_lo(w[j]); a[2 * j + 1] += bf16_hi(w[j]); }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = a[j];
    __syncthreads();
    if (threadIdx.x < ncg * 8) {
      const int c8 = threadIdx.x / 8, j = threadIdx.x % 8;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * ncg + c8][j];
      part[(size_t)chunk * C8 * 8 + (cg0 + c8) * 8 + j] = s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) bias_grad_final_kernel(const float* __restrict__ part, int chunks, int C, float* __restrict__ out, int ld) {
  // 32 channels per block; warp w adds chunks w, w+8, ... (coalesced 128-byte rows), then the 8 partial sums are
  // added in warp order: a fixed summation tree, hence run-to-run identical results
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < C)
    for (int k = w; k < chunks; k += 8) s += part[(size_t)k * C + c];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    out[(size_t)c * ld] = t;
  }
}

static int tgrid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? b : cap);
}

// ---- ae_h term (train_dae.py:317-319): squared_error(h, h_hat).mean() with h_hat = c + h (skip sum) ---------------
// = mean(c^2) of the expanding-path conv c = up_conv_{n_pool+1} (the h parts cancel, also in the gradient): sums2[0] += sum c^2,
// sums2[1] += element count (fp64; data-parallel ranks all-reduce both), then g += 2 c / sums2[1].
__global__ void __launch_bounds__(256) sq_sum_kernel(const uint4* __restrict__ x, long long n8, double* __restrict__ sums2) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    const uint4 v = ldg_nc_v4(x + i);
    const float a0 = bf16_lo(v.x), a1 = bf16_hi(v.x), a2 = bf16_lo(v.y), a3 = bf16_hi(v.y);
    const float a4 = bf16_lo(v.z), a5 = bf16_hi(v.z), a6 = bf16_lo(v.w), a7 = bf16_hi(v.w);
    acc += (double)(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3) + (double)(a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7);
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += red[w];
    atomicAdd(sums2, a);
    if (blockIdx.x == 0) atomicAdd(sums2 + 1, (double)(n8 * 8));
  }
}

// Image 0 of a batched tensor copied to images 1..n-1 (16-byte words): one read, n - 1 writes per thread.
__global__ void __launch_bounds__(256) broadcast_image_kernel(uint4* __restrict__ t, long long words, int n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= words) return;
  const uint4 v = t[i];
  for (int k = 1; k < n; ++k) stg_v4(t + (long long)k * words + i, v);
}

__global__ void __launch_bounds__(256) add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 x = ldg_nc_v4(a + i), y = ldg_nc_v4(b + i);
  stg_v4(out + i, make_uint4(pack_bf16x2(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x)),
                             pack_bf16x2(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y)),
                             pack_bf16x2(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z)),
                             pack_bf16x2(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w))));
}

__global__ void __launch_bounds__(256) ae_grad_add_kernel(uint4* __restrict__ g, const uint4* __restrict__ c, long long n8,
                                                          const double* __restrict__ sums2) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  const float k = (float)(2.0 / sums2[1]);
  const uint4 a = g[i], b = ldg_nc_v4(c + i);
  stg_v4(g + i, make_uint4(pack_bf16x2(bf16_lo(a.x) + k * bf16_lo(b.x), bf16_hi(a.x) + k * bf16_hi(b.x)),
                           pack_bf16x2(bf16_lo(a.y) + k * bf16_lo(b.y), bf16_hi(a.y) + k * bf16_hi(b.y)),
                           pack_bf16x2(bf16_lo(a.z) + k * bf16_lo(b.z), bf16_hi(a.z) + k * bf16_hi(b.z)),
                           pack_bf16x2(bf16_lo(a.w) + k * bf16_lo(b.w), bf16_hi(a.w) + k * bf16_hi(b.w))));
}

}  // namespace iiseg

extern "C" int iiseg_noise_pack(const float* y, const float* noise, float sigma, void* dst, int N, int C, int H, int W,
                                int Cpad, int split, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(y && dst, "noise_pack: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C && Cpad % 8 == 0, "noise_pack: bad shape");
  const long long total = (long long)N * (Cpad / 8) * H * W;
  noise_pack_kernel<<<tgrid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, noise, sigma, reinterpret_cast<uint4*>(dst), C,
                                                                                      H * W, Cpad / 8, total, split);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_loss_grad_terms(const float* logits, const float* target, int N, int C, int H, int W, float lmb, int terms,
                                     double* sums, void* dlogits, int passes, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(logits && target && sums && dlogits, "loss_grad: null tensor");
  IISEG_CHECK(N > 0 && C >= 1 && C <= 16 && H > 0 && W > 0, "loss_grad: bad shape");
  IISEG_CHECK(terms > 0 && terms < 8 && (!(terms & kTermDice) || C >= 2), "loss_grad: terms = bit 0 crossentropy | bit 1 squared_error | bit 2 dice (channel 1)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (passes & 1) IISEG_CUDA(cudaMemsetAsync(sums, 0, ((terms & kTermDice) ? 8 : 4) * sizeof(double), s));
  dim3 grid((H * W + 255) / 256, N);
  for (int pass = 0; pass < 2; ++pass) {
    if (!(passes & (1 << pass))) continue;
    loss_grad_kernel<<<grid, 256, 0, s>>>(logits, target, C, H * W, lmb, sums, reinterpret_cast<__nv_bfloat16*>(dlogits), pass, terms);
    IISEG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int iiseg_loss_grad(const float* logits, const float* target, int N, int C, int H, int W, float lmb, double* sums,
                               void* dlogits, int passes, void* stream) {
  return iiseg_loss_grad_terms(logits, target, N, C, H, W, lmb, iiseg::kTermCE | iiseg::kTermMSE, sums, dlogits, passes, stream);
}

extern "C" int iiseg_sq_sum(const void* x, long long n, double* sums2, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && sums2, "sq_sum: null tensor");
  IISEG_CHECK(n > 0 && n % 8 == 0, "sq_sum: element count must be a positive multiple of 8");
  const long long n8 = n / 8;
  const long long want = (n8 + 255) / 256;
  const int blocks = (int)(want < 4 * 148 ? want : 4 * 148);          // grid-stride: a few blocks per SM, one atomic each
  sq_sum_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(x), n8, sums2);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_broadcast_image(void* t, long long image_bytes, int n, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(t != nullptr, "broadcast_image: null tensor");
  IISEG_CHECK(image_bytes > 0 && image_bytes % 16 == 0 && n >= 1, "broadcast_image: image size must be a positive multiple of 16 bytes");
  if (n == 1) return 0;
  const long long words = image_bytes / 16;
  broadcast_image_kernel<<<(unsigned)((words + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint4*>(t), words, n);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(a && b && out, "add_bf16: null tensor");
  IISEG_CHECK(n > 0 && n % 8 == 0, "add_bf16: element count must be a positive multiple of 8");
  const long long n8 = n / 8;
  add_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), n8);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_ae_grad_add(void* g, const void* c, long long n, const double* sums2, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(g && c && sums2, "ae_grad_add: null tensor");
  IISEG_CHECK(n > 0 && n % 8 == 0, "ae_grad_add: element count must be a positive multiple of 8");
  const long long n8 = n / 8;
  ae_grad_add_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(g), reinterpret_cast<const uint4*>(c), n8, sums2);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_depool2_bwd(const void* gv, const uint32_t* mask, void* gu, int N, int H, int W, int C, int VH, int VW,
                                 int v_h0, int v_w0, int UH, int UW, int u_h0, int u_w0, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(gv && mask && gu, "depool2_bwd: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && C % 8 == 0 && UH >= 1 && UW >= 1 && UH <= 65535 && N <= 65535, "depool2_bwd: bad shape");
  DepoolBwdParams p;
  p.gv = reinterpret_cast<const uint4*>(gv); p.mask = mask; p.gu = reinterpret_cast<uint4*>(gu);
  p.C8 = C / 8; p.H2 = H / 2; p.W2 = W / 2; p.VH = VH; p.VW = VW; p.v_h0 = v_h0; p.v_w0 = v_w0;
  p.UH = UH; p.UW = UW; p.u_h0 = u_h0; p.u_w0 = u_w0;
  dim3 grid(ceil_div(UW * p.C8, 256), UH, N);
  depool_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_pool2_relu_bwd(const void* gpool, const void* pooled, const uint32_t* mask, const uint32_t* zmask, void* ga,
                                    int N, int H, int W, int C, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(gpool && pooled && mask && ga, "pool2_relu_bwd: null tensor");
  IISEG_CHECK(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0 && (H + 1) / 2 <= 65535 && N <= 65535, "pool2_relu_bwd: bad shape");
  const int PW2 = (W + 1) / 2;
  dim3 grid(ceil_div(PW2 * (C / 8), 256), (H + 1) / 2, N);
  pool_relu_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(gpool), reinterpret_cast<const uint4*>(pooled), mask, zmask, reinterpret_cast<uint4*>(ga), H, W, C / 8,
      PW2);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_transpose_shift(const void* x, int N, int H, int W, int Cs, int c0, int C, int h0, int w0, int OH, int OW,
                                     int dh, int dw, void* out, long long ldo, long long row0, int nshift, long long shift_rows,
                                     void* stream) {
  using namespace iiseg;
  IISEG_CHECK(x && out, "transpose_shift: null tensor");
  IISEG_CHECK(N > 0 && C > 0 && c0 >= 0 && c0 + C <= Cs && OH > 0 && OW > 0 && ldo >= (long long)N * OH * OW, "transpose_shift: bad shape");
  IISEG_CHECK(C % 8 == 0 && c0 % 8 == 0 && Cs % 8 == 0 && ldo % 64 == 0, "transpose_shift: channels must come in groups of 8 and ldo in multiples of 64 (16-byte accesses)");
  TransposeParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.N = N; p.H = H; p.W = W; p.Cs = Cs; p.c0 = c0; p.C = C; p.h0 = h0; p.w0 = w0; p.OH = OH; p.OW = OW; p.dh = dh; p.dw = dw;
  IISEG_CHECK(nshift >= 1 && nshift <= 8 && (nshift == 1 || shift_rows >= C), "transpose_shift: bad shift copies");
  p.ldo = ldo; p.row0 = row0; p.P = (long long)N * OH * OW; p.nshift = nshift; p.shift_rows = shift_rows;
  const long long pblocks = (ldo + 63) / 64;
  IISEG_CHECK(pblocks < (1LL << 31) && (C + 63) / 64 <= 65535, "transpose_shift: too large");
  dim3 grid((unsigned)pblocks, (C + 63) / 64);
  transpose_shift_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

static int optimiser_pack(float* w, float* acc, float* b, float* acc_b, const float* g, void* wb, void* wt, int Cout,
                          int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0, int Ci_t, int Co_pad, float lr,
                          float rho, float eps, float* m, float* m_b, const float* a_t, float beta1, void* stream);

extern "C" int iiseg_rmsprop_pack(float* w, float* acc, float* b, float* acc_b, const float* g, void* wb, void* wt, int Cout,
                                  int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0, int Ci_t, int Co_pad, float lr,
                                  float rho, float eps, void* stream) {
  return optimiser_pack(w, acc, b, acc_b, g, wb, wt, Cout, taps, Cin_pad, ldg, g_rstride, bias_col, ci0, Ci_t, Co_pad, lr, rho, eps,
                        nullptr, nullptr, nullptr, 0.f, stream);
}

extern "C" int iiseg_adam_pack(float* w, float* m, float* v, float* b, float* m_b, float* v_b, const float* g, void* wb, void* wt,
                               int Cout, int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0, int Ci_t, int Co_pad,
                               const float* a_t, float beta1, float beta2, float eps, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(m && m_b && a_t, "adam_pack: null tensor");
  return optimiser_pack(w, v, b, v_b, g, wb, wt, Cout, taps, Cin_pad, ldg, g_rstride, bias_col, ci0, Ci_t, Co_pad, 0.f, beta2, eps,
                        m, m_b, a_t, beta1, stream);
}

extern "C" int iiseg_adam_advance(float* state, float lr, float beta1, float beta2, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(state != nullptr, "adam_advance: null state");
  adam_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state, lr, beta1, beta2);
  IISEG_LAUNCH_CHECK();
  return 0;
}

static int optimiser_pack(float* w, float* acc, float* b, float* acc_b, const float* g, void* wb, void* wt, int Cout,
                          int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0, int Ci_t, int Co_pad, float lr,
                          float rho, float eps, float* m, float* m_b, const float* a_t, float beta1, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(w && acc && b && acc_b && g && wb, "rmsprop_pack: null tensor");
  if (g_rstride == 0) g_rstride = 3 * Cin_pad;
  IISEG_CHECK(Cout > 0 && taps > 0 && taps % 3 == 0 && Cin_pad > 0 && g_rstride >= 3 * Cin_pad && g_rstride % 4 == 0 &&
              ldg >= (taps / 3) * g_rstride && bias_col < ldg, "rmsprop_pack: bad shape");
  IISEG_CHECK(wt == nullptr || (ci0 >= 0 && ci0 + Ci_t <= Cin_pad && Co_pad >= Cout), "rmsprop_pack: bad transposed-bank range");
  RmspropParams p;
  p.w = w; p.acc = acc; p.b = b; p.acc_b = acc_b; p.g = g; p.wb = reinterpret_cast<__nv_bfloat16*>(wb); p.wt = reinterpret_cast<__nv_bfloat16*>(wt);
  p.Cout = Cout; p.taps = taps; p.Cin_pad = Cin_pad; p.ldg = ldg; p.bias_col = bias_col; p.ci0 = ci0; p.Ci_t = Ci_t; p.Co_pad = Co_pad; p.g_rstride = g_rstride;
  p.lr = lr; p.rho = rho; p.eps = eps; p.total = (long long)Cout * taps * Cin_pad;
  p.m = m; p.m_b = m_b; p.a_t = a_t; p.beta1 = beta1;
  IISEG_CHECK(Cin_pad % 16 == 0 && Cout % 4 == 0 && ldg % 4 == 0 && (wt == nullptr || (ci0 % 16 == 0 && Ci_t % 16 == 0 && Co_pad % 4 == 0)),
              "rmsprop_pack: channel counts must be padded (Cin %% 16, Cout %% 4, transposed range %% 16)");
  IISEG_CHECK((Cout + 63) / 64 <= 65535, "rmsprop_pack: too many output channels");
  dim3 grid(taps * (Cin_pad / 16), (Cout + 63) / 64);
  rmsprop_pack_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_sum_slabs(const float* in, float* out, int S, long long n, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(in && out && S >= 1 && n > 0 && n % 4 == 0, "sum_slabs: bad arguments");
  sum_slabs_kernel<<<tgrid(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(in),
                                                                                    reinterpret_cast<float4*>(out), S, n / 4);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_bias_grad(const void* g, long long P, int C, float* scratch, int chunks, float* out, int ld, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(g && scratch && out && P > 0 && C > 0 && C % 8 == 0 && chunks >= 1 && chunks <= 1024 && ld >= 1, "bias_grad: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  bias_grad_partial_kernel<<<chunks, 256, 0, s>>>(reinterpret_cast<const uint4*>(g), P, C / 8, scratch);
  IISEG_LAUNCH_CHECK();
  bias_grad_final_kernel<<<(C + 31) / 32, 256, 0, s>>>(scratch, chunks, C, out, ld);
  IISEG_LAUNCH_CHECK();
  return 0;
}

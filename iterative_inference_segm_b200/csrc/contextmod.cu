// Context-module DAE (models/contextmod_dae.py:19-138): a chain of 3x3 convolutions on n_classes (<= 16)
// channels at full image resolution -- conv1 ('same', on [h | y]), PadLayer(32), DilatedConv2DLayer with
// dilation 1, 2, 4, 8, 16, 1 ('valid', rectify), a 1x1 linear conv, channel softmax.
//
// 11 -> 11 channels is no tensor-core shape (N = K = 16 would run the UMMA pipe at a few per cent while the
// operands still cross HBM); the layers are fp32 FMA work on the CUDA cores, exact in the reference's float32:
//   * activations are planar fp32 [N, C, H, W] (the layout of the loop's master y): a warp reads / writes 128
//     contiguous bytes per channel plane, any dilation;
//   * a thread owns two pixels (ow, ow + 32) x all COUT channels in registers; the COUT x CIN x 9 weights travel
//     in the kernel's parameter block (constant bank), so every FFMA takes its weight as a constant operand and
//     the inner loop is 2 loads per 2*COUT FMAs;
//   * conv1 stores straight into the interior of the zero-bordered PadLayer buffer, adding the hoisted
//     iteration-invariant W_h * h term; the last dilated conv carries the 1x1 conv and writes the fp32 NHWC16
//     logits rows iiseg_softmax_update consumes.
// DilatedConv2DLayer semantics (lasagne/layers/conv.py, W of shape (Cin, Cout, 3, 3), unflipped):
//     out[n, f, i, j] = b[f] + sum_{c, r, s} W[c, f, r, s] * in[n, c, i + r*d, j + s*d]
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

template <int CIN, int COUT>
struct CtxParams {
  const float* in; float* out; const float* addend; const int32_t* active;
  int cin, cout;                             // real channel counts (<= CIN / COUT; the rest of w is zero)
  int Hin, Win, in_h0, in_w0, check, dil;
  int Hout, Wout, out_h0, out_w0, OH, OW;
  int relu;
  float b[COUT];
  float w[CIN * 9 * COUT];                   // [ci][tap][co]
  float b2[16];
  float w2[COUT * 16];                       // [ci][co2] of the fused 1x1 tail
};

constexpr int kCtxThreads = 128;             // 4 warps = 4 output rows of 64 pixels

template <int CIN, int COUT, bool kTail>
__global__ void __launch_bounds__(kCtxThreads) ctx_conv_kernel(const __grid_constant__ CtxParams<CIN, COUT> P) {
  const int n = blockIdx.z;
  if (P.active != nullptr && P.active[n] == 0) return;        // frozen image
  const int lane = threadIdx.x & 31;
  const int oh = blockIdx.y * (kCtxThreads / 32) + (threadIdx.x >> 5);
  if (oh >= P.OH) return;
  const int ow0 = blockIdx.x * 64 + lane;
  const bool v0 = ow0 < P.OW, v1 = ow0 + 32 < P.OW;
  float a0[COUT], a1[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) { a0[co] = P.b[co]; a1[co] = P.b[co]; }
  const size_t plane = static_cast<size_t>(P.Hin) * P.Win;
  const float* inb = P.in + static_cast<size_t>(n) * P.cin * plane;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh + P.in_h0 + r * P.dil;
    const bool rok = !P.check || (ih >= 0 && ih < P.Hin);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int iw = ow0 + P.in_w0 + s * P.dil;
      const bool ok0 = v0 && rok && (!P.check || (iw >= 0 && iw < P.Win));
      const bool ok1 = v1 && rok && (!P.check || (iw + 32 >= 0 && iw + 32 < P.Win));
      const float* q = inb + static_cast<ptrdiff_t>(ih) * P.Win + iw;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        if (CIN != 16 || ci < P.cin) {          // (the padded <16,16> instantiation serves any smaller channel count)
          const float x0 = ok0 ? __ldg(q + ci * plane) : 0.f;
          const float x1 = ok1 ? __ldg(q + ci * plane + 32) : 0.f;
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float wv = P.w[(ci * 9 + r * 3 + s) * COUT + co];
            a0[co] = fmaf(x0, wv, a0[co]);
            a1[co] = fmaf(x1, wv, a1[co]);
          }
        }
      }
    }
  }
  if constexpr (!kTail) {
    const size_t oplane = static_cast<size_t>(P.Hout) * P.Wout;
    float* ob = P.out + static_cast<size_t>(n) * P.cout * oplane + static_cast<size_t>(oh + P.out_h0) * P.Wout + ow0 + P.out_w0;
    const float* ab = P.addend == nullptr ? nullptr
                                          : P.addend + (static_cast<size_t>(n) * P.cout * P.OH + oh) * P.OW + ow0;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      if (COUT != 16 || co < P.cout) {
        float r0 = a0[co], r1 = a1[co];
        if (ab != nullptr) {
          if (v0) r0 += __ldg(ab + static_cast<size_t>(co) * P.OH * P.OW);
          if (v1) r1 += __ldg(ab + static_cast<size_t>(co) * P.OH * P.OW + 32);
        }
        if (P.relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); }
        if (v0) ob[co * oplane] = r0;
        if (v1) ob[co * oplane + 32] = r1;
      }
    }
  } else {
    // rectify, then the 1x1 conv (dilconv7, linear) on registers; one fp32 NHWC16 logits row per pixel
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      float* a = px == 0 ? a0 : a1;
      float l[16];
#pragma unroll
      for (int c2 = 0; c2 < 16; ++c2) l[c2] = P.b2[c2];
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        const float x = P.relu ? fmaxf(a[co], 0.f) : a[co];
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) l[c2] = fmaf(x, P.w2[co * 16 + c2], l[c2]);
      }
      if (px == 0 ? v0 : v1) {
        float* o = P.out + ((static_cast<size_t>(n) * P.OH + oh) * P.OW + ow0 + 32 * px) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg_v4(o + 4 * j, make_uint4(__float_as_uint(l[4 * j]), __float_as_uint(l[4 * j + 1]), __float_as_uint(l[4 * j + 2]),
                                       __float_as_uint(l[4 * j + 3])));
      }
    }
  }
}

template <int CIN, int COUT>
static int launch_ctx(const iiseg_ctx_conv_desc* d) {
  static thread_local CtxParams<CIN, COUT> P;      // ~10 KB: filled here, passed to the kernel by value (constant bank)
  static_assert(sizeof(CtxParams<CIN, COUT>) < 32000, "kernel parameter block");
  P.in = d->in; P.out = d->out; P.addend = d->addend; P.active = d->active;
  P.cin = d->Cin; P.cout = d->Cout;
  P.Hin = d->Hin; P.Win = d->Win; P.in_h0 = d->in_h0; P.in_w0 = d->in_w0; P.check = d->check; P.dil = d->dil;
  P.Hout = d->Hout; P.Wout = d->Wout; P.out_h0 = d->out_h0; P.out_w0 = d->out_w0; P.OH = d->OH; P.OW = d->OW;
  P.relu = d->relu;
  for (int co = 0; co < COUT; ++co) P.b[co] = co < d->Cout ? d->bias[co] : 0.f;
  for (int ci = 0; ci < CIN; ++ci)
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < COUT; ++co)
        P.w[(ci * 9 + t) * COUT + co] = (ci < d->Cin && co < d->Cout) ? d->weight[(static_cast<size_t>(ci) * 9 + t) * d->Cout + co] : 0.f;
  const bool tail = d->weight2 != nullptr;
  if (tail) {
    for (int c2 = 0; c2 < 16; ++c2) P.b2[c2] = c2 < d->C2 ? d->bias2[c2] : 0.f;
    for (int co = 0; co < COUT; ++co)
      for (int c2 = 0; c2 < 16; ++c2)
        P.w2[co * 16 + c2] = (co < d->Cout && c2 < d->C2) ? d->weight2[co * d->C2 + c2] : 0.f;
  }
  const dim3 grid(ceil_div(d->OW, 64), ceil_div(d->OH, kCtxThreads / 32), d->N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(d->stream);
  if (tail) ctx_conv_kernel<CIN, COUT, true><<<grid, kCtxThreads, 0, st>>>(P);
  else ctx_conv_kernel<CIN, COUT, false><<<grid, kCtxThreads, 0, st>>>(P);
  IISEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace iiseg

extern "C" int iiseg_ctx_conv_desc_size(void) { return static_cast<int>(sizeof(iiseg_ctx_conv_desc)); }

extern "C" int iiseg_ctx_conv(const iiseg_ctx_conv_desc* d) {
  using namespace iiseg;
  IISEG_CHECK(d != nullptr && d->in && d->out && d->weight && d->bias, "ctx_conv: null tensor");
  IISEG_CHECK(d->N >= 1 && d->Cin >= 1 && d->Cin <= 16 && d->Cout >= 1 && d->Cout <= 16, "ctx_conv: 1..16 channels (got %d -> %d)", d->Cin, d->Cout);
  IISEG_CHECK(d->dil >= 1 && d->OH >= 1 && d->OW >= 1 && d->Hin >= 1 && d->Win >= 1, "ctx_conv: bad shape");
  IISEG_CHECK(d->N <= 65535 && ceil_div(d->OH, 4) <= 65535, "ctx_conv: grid too large");
  if (!d->check)
    IISEG_CHECK(d->in_h0 >= 0 && d->in_w0 >= 0 && d->OH - 1 + d->in_h0 + 2 * d->dil < d->Hin && d->OW - 1 + d->in_w0 + 2 * d->dil < d->Win,
                "ctx_conv: a 'valid' window must lie inside the input (set check for zero padding)");
  const bool tail = d->weight2 != nullptr;
  if (tail) {
    IISEG_CHECK(d->bias2 != nullptr && d->C2 >= 1 && d->C2 <= 16 && d->addend == nullptr, "ctx_conv: tail needs bias2, 1..16 output channels and no addend");
  } else {
    IISEG_CHECK(d->out_h0 >= 0 && d->out_w0 >= 0 && d->out_h0 + d->OH <= d->Hout && d->out_w0 + d->OW <= d->Wout, "ctx_conv: output window outside the output tensor");
  }
  if (d->Cin == 11 && d->Cout == 11) return launch_ctx<11, 11>(d);
  if (d->Cin == 3 && d->Cout == 11 && !tail) return launch_ctx<3, 11>(d);
  return launch_ctx<16, 16>(d);          // any other 1..16 -> 1..16: zero-padded weights
}

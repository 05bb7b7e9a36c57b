// Context-module DAE (models/contextmod_dae.py:19-138): a chain of 3x3 convolutions on n_classes (<= 16)
// channels at full image resolution -- conv1 ('same', on [h | y]), PadLayer(32), DilatedConv2DLayer with
// dilation 1, 2, 4, 8, 16, 1 ('valid', rectify), a 1x1 linear conv, channel softmax.
//
// 11 -> 11 channels is no tensor-core shape (N = K = 16 would run the UMMA pipe at a few per cent while the
// operands still cross HBM); the layers are fp32 FMA work on the CUDA cores, exact in the reference's float32:
//   * activations are planar fp32 [N, C, H, W] (the layout of the loop's master y): a warp reads / writes 128
//     contiguous bytes per channel plane, any dilation;
//   * a thread owns four pixels (2 rows x columns ow, ow + 32) x all COUT channels in registers; the COUT x CIN x 9
//     weights travel in the kernel's parameter block (constant bank) and the FMAs are FFMA2 on output-channel pairs;
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
  static constexpr int CP = (COUT + 1) & ~1;         // output channels padded to whole FFMA2 pairs
  const float* in; float* out; const float* addend; const int32_t* active;
  int cin, cout;                             // real channel counts (<= CIN / COUT; the rest of w is zero)
  int Hin, Win, in_h0, in_w0, check, dil;
  int Hout, Wout, out_h0, out_w0, OH, OW;
  int relu;
  alignas(16) float b[CP];
  alignas(16) float w[CIN * 9 * CP];         // [ci][tap][co]
  float b2[16];
  float w2[COUT * 16];                       // [ci][co2] of the fused 1x1 tail
};

constexpr int kCtxThreads = 128;             // 4 warps; a warp computes 2 output rows x 64 pixels

// two IEEE fp32 FMAs per issued instruction (FFMA2, sm_100): acc.{x,y} = a.{x,y} * b.{x,y} + acc.{x,y}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long dup_f32(float x) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// The kernel is bound by instruction issue, not by memory (ncu: L2 9-21 %, DRAM < 17 %): with scalar FFMAs two thirds of
// the issue slots are FMAs and the FMA pipe sits at 48 %.  So: a thread owns FOUR pixels (2 rows x columns ow, ow + 32)
// x all channels, accumulators are channel PAIRS and the FMAs are FFMA2 (x duplicated into both halves, two adjacent
// output-channel weights from the constant bank).  Layouts: planar [N,C,H,W] (the loop's master y, the image) costs one
// 4-byte load + its 64-bit address per (pixel, channel, tap); the module's own intermediate tensors are therefore
// channels-last with the channel count padded to a multiple of four, [N,H,W,CP4]: a pixel's channels are CP4/4 16-byte
// loads (three for 11 channels), stores likewise, and a warp still covers one contiguous span per access.
template <int CIN, int COUT, bool kTail, bool kInNHWC, bool kOutNHWC>
__global__ void __launch_bounds__(kCtxThreads) ctx_conv_kernel(const __grid_constant__ CtxParams<CIN, COUT> P) {
  constexpr int CP = CtxParams<CIN, COUT>::CP, NP = CP / 2;
  constexpr int CPI = (CIN + 3) & ~3, CPO = (COUT + 3) & ~3;          // channels-last pitches
  const int n = blockIdx.z;
  if (P.active != nullptr && P.active[n] == 0) return;        // frozen image
  const int lane = threadIdx.x & 31;
  const int oh = 2 * (blockIdx.y * (kCtxThreads / 32) + (threadIdx.x >> 5));
  if (oh >= P.OH) return;
  const int ow0 = blockIdx.x * 64 + lane;
  const bool vr[2] = {true, oh + 1 < P.OH};
  const bool vc[2] = {ow0 < P.OW, ow0 + 32 < P.OW};
  unsigned long long acc[4][NP];             // pixel k = 2 * row + col
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const unsigned long long bj = *reinterpret_cast<const unsigned long long*>(&P.b[2 * j]);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k][j] = bj;
  }
  const size_t plane = static_cast<size_t>(P.Hin) * P.Win;
  // channels-last pitches: compile-time for the exact instantiations, from the real channel counts for the padded <16,16>
  const int cpi = CIN == 16 ? ((P.cin + 3) & ~3) : CPI, cpo = COUT == 16 ? ((P.cout + 3) & ~3) : CPO;
  const float* inb = P.in + static_cast<size_t>(n) * plane * (kInNHWC ? cpi : P.cin);        // warp-uniform
  // The tap loops stay ROLLED: unrolled, the kernel is ~4300 straight-line instructions that every warp runs once, and
  // instruction fetch becomes the top stall (ncu: no_instruction 2.1 warps per issue); one tap is ~500 instructions.
  // The weights of tap (r, s) are then read from the constant bank at a warp-uniform runtime offset.
#pragma unroll 1
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ih = oh + P.in_h0 + r * P.dil, iw = ow0 + P.in_w0 + s * P.dil;
      const float* wt = P.w + (r * 3 + s) * CP;          // + ci * 9 * CP + 2 * j
      bool ok[4];
      uint32_t off[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = ih + (k >> 1), x = iw + 32 * (k & 1);
        // valid pixels of a 'valid' conv read inside the plane by construction; overhanging (never stored) pixels and,
        // with `check`, the taps in the zero padding are redirected to element 0 (check: and contribute zero)
        ok[k] = vr[k >> 1] && vc[k & 1] && (!P.check || (y >= 0 && y < P.Hin && x >= 0 && x < P.Win));
        off[k] = ok[k] ? static_cast<uint32_t>(y * P.Win + x) : 0u;
      }
      if constexpr (kInNHWC) {
        float xv[4][CPI];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4* q = reinterpret_cast<const float4*>(inb + static_cast<size_t>(off[k]) * cpi);
#pragma unroll
          for (int v = 0; v < CPI / 4; ++v) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (CIN != 16 || 4 * v < cpi) t = __ldg(q + v);        // (channels-last inputs are 'valid' convs: no zero padding to emulate)
            xv[k][4 * v] = t.x; xv[k][4 * v + 1] = t.y; xv[k][4 * v + 2] = t.z; xv[k][4 * v + 3] = t.w;
          }
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          if (CIN != 16 || ci < P.cin) {
            unsigned long long xx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) xx[k] = dup_f32(xv[k][ci]);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              const unsigned long long wv = *reinterpret_cast<const unsigned long long*>(wt + ci * 9 * CP + 2 * j);
#pragma unroll
              for (int k = 0; k < 4; ++k) ffma2(acc[k][j], xx[k], wv);
            }
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          if (CIN != 16 || ci < P.cin) {          // (the padded <16,16> instantiation serves any smaller channel count)
            const float* pl = inb + ci * plane;
            unsigned long long xx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x = __ldg(pl + off[k]);
              xx[k] = dup_f32((P.check && !ok[k]) ? 0.f : x);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              const unsigned long long wv = *reinterpret_cast<const unsigned long long*>(wt + ci * 9 * CP + 2 * j);
#pragma unroll
              for (int k = 0; k < 4; ++k) ffma2(acc[k][j], xx[k], wv);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!(vr[k >> 1] && vc[k & 1])) continue;
    const int ohk = oh + (k >> 1), owk = ow0 + 32 * (k & 1);
    float a[CP > CPO ? CP : CPO];
#pragma unroll
    for (int j = 0; j < NP; ++j) { const float2 t = unpack_f32x2(acc[k][j]); a[2 * j] = t.x; a[2 * j + 1] = t.y; }
#pragma unroll
    for (int c = CP; c < CPO; ++c) a[c] = 0.f;
    if constexpr (kTail) {
      // rectify, then the 1x1 conv (dilconv7, linear) on registers; one fp32 NHWC16 logits row per pixel
      float l[16];
#pragma unroll
      for (int c2 = 0; c2 < 16; ++c2) l[c2] = P.b2[c2];
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        const float x = P.relu ? fmaxf(a[co], 0.f) : a[co];
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) l[c2] = fmaf(x, P.w2[co * 16 + c2], l[c2]);
      }
      float* o = P.out + ((static_cast<size_t>(n) * P.OH + ohk) * P.OW + owk) * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        stg_v4(o + 4 * j, make_uint4(__float_as_uint(l[4 * j]), __float_as_uint(l[4 * j + 1]), __float_as_uint(l[4 * j + 2]),
                                     __float_as_uint(l[4 * j + 3])));
    } else if constexpr (kOutNHWC) {
      // channels-last out (and addend): [N,Hout,Wout,CPO]; the pad channels carry zero weights and bias, so they store 0
      float* o = P.out + ((static_cast<size_t>(n) * P.Hout + ohk + P.out_h0) * P.Wout + owk + P.out_w0) * cpo;
      const float4* ab = P.addend == nullptr ? nullptr
                                             : reinterpret_cast<const float4*>(P.addend + ((static_cast<size_t>(n) * P.OH + ohk) * P.OW + owk) * cpo);
#pragma unroll
      for (int v = 0; v < CPO / 4; ++v) {
        if (COUT == 16 && 4 * v >= cpo) break;
        float4 t = make_float4(a[4 * v], a[4 * v + 1], a[4 * v + 2], a[4 * v + 3]);
        if (ab != nullptr) { const float4 u = __ldg(ab + v); t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        if (P.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
        stg_v4(o + 4 * v, make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w)));
      }
    } else {
      const size_t oplane = static_cast<size_t>(P.Hout) * P.Wout;
      float* ob = P.out + static_cast<size_t>(n) * P.cout * oplane + static_cast<size_t>(ohk + P.out_h0) * P.Wout + owk + P.out_w0;
      const float* ab = P.addend == nullptr ? nullptr : P.addend + (static_cast<size_t>(n) * P.cout * P.OH + ohk) * P.OW + owk;
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        if (COUT != 16 || co < P.cout) {
          float v = a[co];
          if (ab != nullptr) v += __ldg(ab + static_cast<size_t>(co) * P.OH * P.OW);
          if (P.relu) v = fmaxf(v, 0.f);
          ob[co * oplane] = v;
        }
      }
    }
  }
}

template <int CIN, int COUT>
static int launch_ctx(const iiseg_ctx_conv_desc* d) {
  using PT = CtxParams<CIN, COUT>;
  constexpr int CP = PT::CP;
  static thread_local PT P;      // ~5-11 KB: filled here, passed to the kernel by value (constant bank)
  static_assert(sizeof(PT) < 32000, "kernel parameter block");
  P.in = d->in; P.out = d->out; P.addend = d->addend; P.active = d->active;
  P.cin = d->Cin; P.cout = d->Cout;
  P.Hin = d->Hin; P.Win = d->Win; P.in_h0 = d->in_h0; P.in_w0 = d->in_w0; P.check = d->check; P.dil = d->dil;
  P.Hout = d->Hout; P.Wout = d->Wout; P.out_h0 = d->out_h0; P.out_w0 = d->out_w0; P.OH = d->OH; P.OW = d->OW;
  P.relu = d->relu;
  for (int co = 0; co < CP; ++co) P.b[co] = co < d->Cout ? d->bias[co] : 0.f;
  for (int ci = 0; ci < CIN; ++ci)
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < CP; ++co)
        P.w[(ci * 9 + t) * CP + co] = (ci < d->Cin && co < d->Cout) ? d->weight[(static_cast<size_t>(ci) * 9 + t) * d->Cout + co] : 0.f;
  const bool tail = d->weight2 != nullptr;
  if (tail) {
    for (int c2 = 0; c2 < 16; ++c2) P.b2[c2] = c2 < d->C2 ? d->bias2[c2] : 0.f;
    for (int co = 0; co < COUT; ++co)
      for (int c2 = 0; c2 < 16; ++c2)
        P.w2[co * 16 + c2] = (co < d->Cout && c2 < d->C2) ? d->weight2[co * d->C2 + c2] : 0.f;
  }
  const dim3 grid(ceil_div(d->OW, 64), ceil_div(d->OH, 2 * (kCtxThreads / 32)), d->N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(d->stream);
  const bool in_cl = d->in_nhwc != 0, out_cl = d->out_nhwc != 0;
#define IISEG_CTX_LAUNCH(T, I, O) ctx_conv_kernel<CIN, COUT, T, I, O><<<grid, kCtxThreads, 0, st>>>(P)
  if constexpr (CIN < 4) {          // (an image: planar in only)
    if (out_cl) IISEG_CTX_LAUNCH(false, false, true); else IISEG_CTX_LAUNCH(false, false, false);
  } else if (tail) {
    if (in_cl) IISEG_CTX_LAUNCH(true, true, false); else IISEG_CTX_LAUNCH(true, false, false);
  } else if (in_cl) {
    if (out_cl) IISEG_CTX_LAUNCH(false, true, true); else IISEG_CTX_LAUNCH(false, true, false);
  } else {
    if (out_cl) IISEG_CTX_LAUNCH(false, false, true); else IISEG_CTX_LAUNCH(false, false, false);
  }
#undef IISEG_CTX_LAUNCH
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
  IISEG_CHECK(d->N <= 65535 && ceil_div(d->OH, 8) <= 65535, "ctx_conv: grid too large");
  if (!d->check)
    IISEG_CHECK(d->in_h0 >= 0 && d->in_w0 >= 0 && d->OH - 1 + d->in_h0 + 2 * d->dil < d->Hin && d->OW - 1 + d->in_w0 + 2 * d->dil < d->Win,
                "ctx_conv: a 'valid' window must lie inside the input (set check for zero padding)");
  const bool tail = d->weight2 != nullptr;
  if (tail) {
    IISEG_CHECK(d->bias2 != nullptr && d->C2 >= 1 && d->C2 <= 16 && d->addend == nullptr, "ctx_conv: tail needs bias2, 1..16 output channels and no addend");
  } else {
    IISEG_CHECK(d->out_h0 >= 0 && d->out_w0 >= 0 && d->out_h0 + d->OH <= d->Hout && d->out_w0 + d->OW <= d->Wout, "ctx_conv: output window outside the output tensor");
  }
  IISEG_CHECK(!(d->in_nhwc && (d->Cin < 4 || d->check)), "ctx_conv: a channels-last input needs >= 4 channels and a 'valid' window (check = 0)");
  if (d->Cin == 11 && d->Cout == 11) return launch_ctx<11, 11>(d);
  if (d->Cin == 3 && d->Cout == 11 && !tail) return launch_ctx<3, 11>(d);
  return launch_ctx<16, 16>(d);          // any other 1..16 -> 1..16: zero-padded weights
}

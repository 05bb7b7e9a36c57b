/*
 * iiseg.h -- C ABI of libiiseg.so: the B200 (sm_100a) kernels behind the
 * iterative-inference hot path of adri-romsor/iterative_inference_segm.
 *
 * The reference has no FFI of its own; its seams are the four compiled Theano
 * callables pred_fcn_fn / pred_dae_fn / de_fn / val_fn
 * (iterative_inference.py:187-210).  Each entry point below replaces the
 * Theano/Lasagne op (cited per function) those callables are built from.  The
 * Python host (iterative_inference_segm_b200/) binds this library with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch allocator);
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on
 *    that stream and safe to capture into a CUDA Graph;
 *  - return value: 0 = OK, negative = error, text via iiseg_last_error();
 *  - activations are NHWC bf16 with the channel count padded (see each call);
 *    the boundary tensors (images, y, targets) are the reference's NCHW fp32;
 *  - no call falls back to the CPU.
 */
#ifndef IISEG_H_
#define IISEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IISEG_ABI_VERSION 5
#define IISEG_MAX_SRC 6
#define IISEG_MAX_WGROUPS 9

/* ---- library ----------------------------------------------------------- */
int iiseg_abi_version(void);
/* sizeof(iiseg_conv_desc) and offsetof(iiseg_conv_desc, upd_cpad) (its last field): lets a binding check its mirror of the
 * struct without a GPU. */
int iiseg_conv_desc_size(void);
int iiseg_conv_desc_last_offset(void);
int iiseg_deconv_desc_size(void);      /* sizeof(iiseg_deconv_desc) */
const char* iiseg_last_error(void);
/* 0 if device `dev` is compute capability 10.x; negative otherwise. */
int iiseg_device_check(int dev);
/* After a failed stream sync: copies the kernels' pinned-host diagnostic words
 * (which pipeline barrier timed out) into out[0..n). Returns words written. */
int iiseg_read_diag(int32_t* out, int n);
/* Tuning aid: with IISEG_CONV_DBG=4 block 0 of every conv launch stamps clock64() at fixed points
 * of its first 32 tiles (16 slots each); returns the words copied. */
int iiseg_debug_read_timeline(long long* out, int n);
/* Leave `n` SMs free: the persistent kernels (one CTA per SM, all of its shared memory) size their grids for SM count - n,
 * so that another stream's kernels -- NCCL's all-reduce CTAs under the data-parallel backward pass -- can be resident at the
 * same time.  Returns the previous value; 0 restores the full machine. */
int iiseg_reserve_sms(int n);
/* Number of kernel launches issued through this library since load. */
int64_t iiseg_launch_count(void);

/* ---- layout conversion at the boundary ---------------------------------- */
/* NCHW fp32 [N,C,H,W] -> NHWC bf16 [N,H,W,Cpad], channels >= C zero-filled.
 * split = 1: dst is [N,H,W,2*Cpad], the (hi | lo) bf16 pair of each fp32 value
 * (hi = bf16(x) in [0,Cpad), lo = bf16(x - hi) in [Cpad,2*Cpad)): the activation
 * format of the fp32-accurate conv variant (iiseg_conv_desc.split). */
int iiseg_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C,
                                     int H, int W, int Cpad, int split, void* stream);
/* NHWC bf16 [N,H,W,Cpad] -> NCHW fp32 [N,C,H,W] (first C channels); split = 1:
 * src is [N,H,W,2*Cpad] and dst = hi + lo. */
int iiseg_unpack_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C,
                                       int H, int W, int Cpad, int split, void* stream);
/* NHWC bf16 [pixels, C] (split = 1: the (hi | lo) pair [pixels, 2*C], dst = hi + lo) -> NHWC fp32 [pixels, C]; C % 8 == 0.
 * Feeds iiseg_channel_stats with a rectified conv output (the batch-statistics BatchNormLayer of DePool2D's mask
 * sub-graph, layers/mylayers.py:91-93 with models/fcn_down.py:113-115). */
int iiseg_widen_nhwc_bf16_to_f32(const void* src, float* dst, long long pixels, int C, int split, void* stream);
/* NHWC fp32 [N,H,W,Cpad] -> NCHW fp32 [N,C,H,W]. */
int iiseg_unpack_nhwc_f32_to_nchw_f32(const float* src, float* dst, int N, int C,
                                      int H, int W, int Cpad, void* stream);

/* ---- convolution: lasagne Conv2DLayer(flip_filters=False) ---------------
 * Replaces every Conv2DLayer on the path (models/fcn_down.py:102-104,
 * models/fcn_up.py:84-86, models/fcn8.py:33-85): stride-1 cross-correlation as
 * an implicit GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators), with
 * bias, optional ReLU and optional skip-sum (ElemwiseSumLayer,
 * models/fcn_up.py:96-102) fused in the epilogue and the channel concat of a
 * second source (ConcatLayer((h, pool4)), models/model_helpers.py:93-94) done
 * in the loader: the K loop walks src0's channels, then src1's. */
typedef struct iiseg_conv_desc {
  /* Up to IISEG_MAX_SRC activation sources concatenated along channels in the loader, in this order
   * (source i supplies C[i] channels; unused entries are NULL / 0 and sources are packed from
   * index 0).  Each is a view of an NHWC bf16 tensor [N,H,W,Cs[i]]: the first C[i] channels from
   * the given pointer, Cs[i] channels per pixel in memory (0 = dense, Cs = C).  C[i] % 64 == 0. */
  const void* src[IISEG_MAX_SRC];
  int C[IISEG_MAX_SRC];
  int Cs[IISEG_MAX_SRC];
  int N, H, W;        /* input extent                                       */
  /* Split-K GEMM views (1x1 launches, the weight-gradient GEMMs of the training step): image n of the batch is the
   * n-th K slab of one row-major matrix.  src_image_stride = elements between consecutive images of the source (0 =
   * dense H*W*Cs); weight_ld = elements per weight row (0 = R*S*sum C); the weight K coordinate of image n starts at
   * n * w_koff.  Each image then yields its own partial product in `out` [N,..]. */
  long long src_image_stride, weight_ld;
  int w_koff;
  /* Weight row groups (1x1 launches): `weight` holds w_rows_total rows; the Cout output channels are w_groups groups
   * of Cout / w_groups channels, and group g reads rows [w_group_row[g], + Cout / w_groups) at K coordinate +
   * w_group_koff[g] (a multiple of 8; reads outside the matrix give zero).  With K = pixel index on a zero-padded
   * grid of pitch Gw (Gw % 8 == 0), the nine taps of a weight gradient
   *   dW[co][r][s][ci] = sum_p g[co][p] * x[ci][p + r*Gw + s]
   * are nine such views of three column-shifted copies of x^T (rows s*Cin.., K offset r*Gw): no 9-fold im2col. */
  int w_groups, w_rows_total;
  int w_group_koff[IISEG_MAX_WGROUPS];
  int w_group_row[IISEG_MAX_WGROUPS];
  /* Fused DePool2D loader (layers/mylayers.py:88-115; 3x3 convs with Cout in {16,64,128} on the halo-tile kernel):
   * depool_mask != NULL makes the conv input the VIRTUAL map v [N,H,W,C] = DePool2D(u, mask), never written to
   * memory.  src[0] is then the pooled tensor u, dense [N,depool_UH,depool_UW,C[0]], whose element (0,0) sits at
   * pooled-grid position (depool_h0, depool_w0); depool_mask is the tie mask [N,H/2,W/2,C[0]/8] of
   * iiseg_maxpool2_mask_fwd; pooled positions outside u read as zero. */
  const uint32_t* depool_mask;
  int depool_UH, depool_UW, depool_h0, depool_w0;
  /* DePool2D fused into the PRODUCER's epilogue: instead of its output u (bf16, Cout % 64 == 0, after bias / skip-sum)
   * the conv writes v = DePool2D(u, depool_out_mask) restricted to a window.  The launch's output pixel (0,0) sits at
   * (depool_out_ph0, depool_out_pw0) of the pooled grid depool_out_H2 x depool_out_W2, over which depool_out_mask is
   * the tie mask [N,H2,W2,Cout/8]; depool_out is dense [N,depool_out_VH,depool_out_VW,Cout], its element (0,0) being
   * full-resolution pixel (depool_out_h0, depool_out_w0).  Every window position under an output pixel is written
   * (zeros where the mask is clear); positions of a trailing odd row / column are never written and must be zero. */
  void* depool_out;
  const uint32_t* depool_out_mask;
  int depool_out_VH, depool_out_VW, depool_out_h0, depool_out_w0;
  int depool_out_H2, depool_out_W2, depool_out_ph0, depool_out_pw0;
  const void* weight; /* bf16 [Cout][R*S][sum C] (K-major GEMM B operand)   */
  /* Optional (3x3, Cout = 16, one 64-channel source, fp32 or fused-update output: the DAE's logits conv): the same filter
   * re-packed for the N-packed kernel, bf16 [39*16][64] -- for filter column s and input-line phase c = j + r (0..5) the
   * 16-row blocks W[.][c - j][s][.] of the output pixels j = max(0, c-2) .. min(3, c) one after the other, in (s, c) order,
   * with three zero blocks behind block (0, 0).  Four vertically adjacent output pixels then share one accumulator row
   * (N = 16..48 per instruction instead of 16).  NULL: the plain kernels run. */
  const void* weight_npack;
  const float* bias;  /* fp32 [Cout]                                        */
  /* Optional per-channel affine applied AFTER the rectifier and before the pool / store: x * post_scale[c] + post_shift[c]
   * (fp32 [Cout] each, both or neither).  Folds the deterministic BatchNormLayer that follows the rectified convs of the
   * DAE's contracting path when bn=1 (models/fcn_down.py:113-115: conv -> rectify -> BatchNormLayer -> Pool2DLayer; with
   * deterministic=True, iterative_inference.py:189-190, the layer is (x - mean) * (gamma * inv_std) + beta on its stored
   * averages).  The fused pool and its tie mask then see the normalised values, as DePool2D does. */
  const float* post_scale;
  const float* post_shift;
  /* Optional third vector (fp32 [Cout]): when non-NULL the epilogue evaluates lasagne's own expression,
   * ((x - post_mean[c]) * post_scale[c]) + post_shift[c] with post_scale = gamma * inv_std and post_shift = beta, one fp32
   * rounding per operation instead of one fused multiply-add: values that round onto the window's exact zeros then tie in
   * the pool exactly where they do in the reference (DePool2D's tie-inclusive mask). */
  const float* post_mean;
  int Cout;           /* padded: 16, or a multiple of 64                    */
  int R, S, pad;      /* filter extent and symmetric zero padding           */
  /* Output window: out pixel (oh,ow) is conv pixel (oh+oh0, ow+ow0); only the
   * OH x OW window is computed (centre crop of up_conv1, CroppingLayer,
   * layers/mylayers.py:36-57).                                             */
  int oh0, ow0, OH, OW;
  void* out;          /* NHWC [N,OH,OW,Cout], bf16 or fp32 per out_f32      */
  /* out_stride > 1 (the output phases of a transposed convolution, lasagne Deconv2DLayer(4, stride=2) of the DAE's
   * unpool_type='standard', models/fcn_up.py:37-63, run as four 2x2 convolutions): output pixel (oh, ow) of this launch is
   * stored at pixel (oh*out_stride + out_h0, ow*out_stride + out_w0) of the tensor `out` [N,out_H,out_W,Cout]; the skip-sum
   * operand is read with the same stride: addend[oh*out_stride + ah0, ow*out_stride + aw0].  Plain stores only (no
   * pool / update fusion); per-tap and CTA-pair kernels. */
  int out_stride, out_H, out_W, out_h0, out_w0;
  const void* addend; /* NHWC bf16 [N,AH,AW,Cout]; addend[oh+ah0, ow+aw0] is added
                       * before the store (skip-sum with a cropped partner), or NULL */
  int AH, AW, ah0, aw0;
  /* addend_f32 = 1: `addend` is fp32 [N,AH,AW,Cout] and is added before the ReLU in fp32 (Cout %
   * 64 == 0).  Used to hoist the iteration-invariant half of a concat conv out of the loop:
   * conv(concat(h, x)) = conv_h(h) + b (computed once, out_f32) + conv_x(x) (every iteration). */
  int addend_f32;
  /* addend_cs > 0: channels per pixel of the `addend` tensor in memory (default: Cout, or 2*Cout with split); the
   * first Cout channels from the given pointer are added.  The mixed-precision DAE (fp32-accurate contracting path,
   * bf16 expanding path) adds the hi halves of the contracting path's (hi | lo) pool tensors this way. */
  int addend_cs;
  /* Fused Pool2DLayer(2) (+ DePool2D mask): when `pooled` != NULL the conv output is max-pooled
   * 2x2/stride 2 (floor) in the epilogue and only `pooled` [N,OH/2,OW/2,Cout] bf16 and, if
   * non-NULL, `pool_mask` [N,OH/2,OW/2,Cout/8] (nibble layout of iiseg_maxpool2_mask_fwd) are
   * written; `out` is not touched.  Needs Cout % 64 == 0.                                  */
  void* pooled;
  uint32_t* pool_mask;
  /* training only (may be NULL): same layout as pool_mask, bit set iff the pre-rectifier value of that element is
   * exactly 0 -- lasagne's rectify is 0.5*(x+|x|), whose gradient at 0 is 0.5 (iiseg_pool2_relu_bwd). */
  uint32_t* pool_zmask;
  /* pool_H > 0: `pooled`/`pool_mask` are full [N,pool_H,pool_W,..] tensors and this launch writes
   * only the pooled window of its (even-aligned) output window: rows [oh0/2, oh0/2 + OH/2).
   * pool_H == 0: they are dense [N,OH/2,OW/2,..].                                            */
  int pool_H, pool_W;
  int relu;           /* 1: rectify (Lasagne default), 0: linear            */
  /* split = 1 (fp32-accurate variant): out / addend / pooled carry 2*Cout channels per pixel, the
   * bf16 pair hi = bf16(x) in [0,Cout) and lo = bf16(x - hi) in [Cout,2*Cout) of the fp32 result;
   * pool and tie mask compare hi+lo.  The caller feeds (hi | lo | hi) activation sources against
   * (W_hi | W_hi | W_lo) weights, so the GEMM accumulates hi*hi + lo*hi + hi*lo in fp32.        */
  int split;
  int out_f32;        /* 1: fp32 output [N,OH,OW,Cout] (no pool, no split)      */
  int out_cs;         /* fp32 outputs: channels per pixel of the destination (0 = Cout): the conv writes
                       * Cout channels at `out` inside a wider tensor (DenseNet stack, ConcatLayer-free) */
  /* Fused softmax tail + iterative-inference update (Cout == 16, the DAE's last conv up_conv1):
   * when upd_y != NULL the logits are not stored; for every image n with upd_active[n] != 0 (NULL =
   * all) the epilogue does  p = softmax over the first upd_C channels;  g = y - p;
   * y <- clip(y - upd_step*g, 0, 1)  on the fp32 NCHW master upd_y [N,upd_C,OH,OW] in place, writes
   * the bf16 NHWC copy upd_y_bf16 [N,OH,OW,upd_cpad], and adds sum_pixels ||g||_2 of image n, in
   * 2^-40 fixed point, to upd_norm_acc[n] (models/fcn_up.py:154-169, iterative_inference.py:267-277;
   * same arithmetic as iiseg_softmax_update).  iiseg_norm_finalize_fixed consumes upd_norm_acc. */
  float* upd_y;
  void* upd_y_bf16;
  const int32_t* upd_active;
  uint64_t* upd_norm_acc;
  float upd_step;
  /* upd_step_dev != NULL: the step is read from device memory at run time instead (one captured CUDA graph then
   * serves every step value of the iterative_inference_valid.py sweep). */
  const float* upd_step_dev;
  /* upd_split = 1: upd_y_bf16 is [N,OH,OW,2*upd_cpad], the (hi | lo) bf16 pair of the updated y (the first conv of
   * the next iteration is an fp32-accurate `split` conv). */
  int upd_C, upd_split, upd_cpad;
} iiseg_conv_desc;
int iiseg_conv2d_fwd(const iiseg_conv_desc* d, void* stream);
/* Which kernel the calling thread's last iiseg_conv2d_fwd launched: *kernel = 0 per-tap, 1 CTA pair (cta_group::2),
 * 2 halo tile, 3 N-packed 16-channel kernel; *bn = N tile, *kb = channels per K block.  Measurement tooling (bench.py groups launches by kernel). */
int iiseg_last_conv_plan(int* kernel, int* bn, int* kb);

/* ---- pooling: lasagne Pool2DLayer(2) + DePool2D -------------------------
 * 2x2/stride-2 max, floor (models/fcn_down.py:122).  `mask` (may be NULL)
 * receives the tie-inclusive argmax mask DePool2D derives with
 * T.grad(pool, ones) (layers/mylayers.py:111-112): four bits per (window,
 * channel), one per window position pos = 2*dy+dx, set iff x[2oh+dy, 2ow+dx] ==
 * window max; eight channels per uint32, layout [N,H/2,W/2,C/8]; inside a word
 * channel j of the group sits at bit 16*(j&1) + 4*(j>>1) + pos.  C multiple of 8. */
int iiseg_maxpool2_mask_fwd(const void* x, void* pooled, uint32_t* mask, int N,
                            int H, int W, int C, void* stream);
/* DePool2D (layers/mylayers.py:88-115): out[2oh+dy,2ow+dx] = u[oh,ow] where
 * the mask bit is set, else 0; trailing odd row/col of the HxW output = 0. */
int iiseg_unpool2_mask_fwd(const void* u, const uint32_t* mask, void* out, int N,
                           int H, int W, int C, void* stream);
/* Windowed form: only the output window [o_h0,o_h0+OH) x [o_w0,o_w0+OW) of the HxW map is
 * produced (dense [N,OH,OW,C]); `u` is a dense [N,UH,UW,C] window of the pooled map whose
 * element (0,0) is pooled position (u_h0,u_w0).  The expanding path only ever needs the
 * dependency cone of the final centre crop (CroppingLayer, models/fcn_up.py:106-113).
 * split = 1: `u` and `out` carry the (hi | lo) bf16 pair of an fp32 map, 2*C channels per pixel
 * (see iiseg_conv_desc.split); the mask still has C channels and gates both halves.
 * split = 2: `u` is such a pair tensor (2*C channels per pixel) but `out` is plain bf16 [N,OH,OW,C]: the hi halves
 * gated by the mask (mixed precision: fp32-accurate contracting path, bf16 expanding path). */
int iiseg_unpool2_mask_window_fwd(const void* u, const uint32_t* mask, void* out, int N,
                                  int H, int W, int C, int UH, int UW, int u_h0, int u_w0,
                                  int OH, int OW, int o_h0, int o_w0, int split, void* stream);

/* ---- transposed convolution: lasagne Deconv2DLayer ----------------------
 * models/fcn8.py:90-91,100-101,109-110 (crop='valid', flip_filters=False,
 * linear) on <=16-channel score maps, fp32 NHWC16 in and out.  Computes the
 * out window [oh0,oh0+OH) x [ow0,ow0+OW) of the (H-1)*stride+k output, adds
 * bias and, if `addend` != NULL, addend[(oh+ah0),(ow+aw0)] (ElemwiseSumLayer
 * with centre cropping, models/fcn8.py:94-97).  weight fp32 [k][k][16][16] =
 * Wt[a][b][ci][co], already flipped/transposed by the host packer. */
typedef struct iiseg_deconv_desc {
  const float* x; int N, H, W;       /* NHWC16 fp32 input                  */
  const float* weight; const float* bias; int k, stride;
  int oh0, ow0, OH, OW;
  const float* addend; int AH, AW, ah0, aw0;  /* NHWC16 fp32 [N,AH,AW,16]  */
  float* out;                        /* NHWC16 fp32 [N,OH,OW,16]           */
} iiseg_deconv_desc;
int iiseg_deconv2d_fwd(const iiseg_deconv_desc* d, void* stream);

/* ---- context-module DAE: small-channel (dilated) 3x3 convolution --------
 * models/contextmod_dae.py:72-103 (kind='contextmod'): Conv2DLayer(n_classes, 3, pad='same', flip_filters=False) on
 * [h | y], PadLayer(32), DilatedConv2DLayer(n_classes, 3, dilation 1/2/4/8/16/1, 'valid', rectify) and a 1x1 linear
 * DilatedConv2DLayer, all on <= 16 channels at image resolution: fp32 FMA work, exact float32 like the reference.
 * Planar fp32 tensors.  For n < N, f < Cout, (oh, ow) in [0,OH) x [0,OW):
 *   v = bias[f] + sum_{c,r,s} weight[c][r][s][f] * in[n][c][oh + in_h0 + r*dil][ow + in_w0 + s*dil]
 *       (check != 0: taps outside [0,Hin) x [0,Win) read zero -- 'same' padding; check == 0: they must be inside)
 *   v += addend[n][f][oh][ow]            (addend != NULL: planar [N,Cout,OH,OW], the hoisted W_h * h term)
 *   out[n][f][oh + out_h0][ow + out_w0] = relu ? max(v, 0) : v          (out planar [N,Cout,Hout,Wout])
 * weight2 != NULL fuses the 1x1 linear conv behind it: out is then fp32 NHWC16 [N,OH,OW,16] logits rows,
 *   out[n][oh][ow][g] = bias2[g] + sum_f weight2[f][g] * (relu ? max(v_f, 0) : v_f), zero for g >= C2.
 * weight / bias / weight2 / bias2 are HOST pointers (fp32): they are copied into the kernel's parameter block at
 * launch (and with it into a captured CUDA graph).  Images with active[n] == 0 are skipped (active may be NULL). */
typedef struct iiseg_ctx_conv_desc {
  const float* in; int N, Cin, Hin, Win;
  int in_h0, in_w0, check, dil;
  float* out; int Cout, Hout, Wout, out_h0, out_w0, OH, OW;
  const float* weight;               /* HOST [Cin][3][3][Cout]              */
  const float* bias;                 /* HOST [Cout]                         */
  const float* addend;
  const int32_t* active;
  int relu;
  const float* weight2;              /* HOST [Cout][C2] or NULL             */
  const float* bias2;                /* HOST [C2]                           */
  int C2;
  int in_nhwc, out_nhwc;             /* != 0: `in` is channels-last [N,Hin,Win,(Cin+3)&~3] / `out` and `addend` are
                                        [N,Hout,Wout,(Cout+3)&~3] and [N,OH,OW,(Cout+3)&~3] (pad channels read as anything
                                        with zero weight, written as 0); the layout of the module's intermediate tensors.
                                        in_nhwc requires check == 0 and Cin >= 4  */
  void* stream;
} iiseg_ctx_conv_desc;
int iiseg_ctx_conv_desc_size(void);
int iiseg_ctx_conv(const iiseg_ctx_conv_desc* d);

/* ---- softmax tail + iterative-inference update --------------------------
 * Channel softmax (models/fcn_up.py:154-169, models/fcn8.py:120-191) of fp32
 * NHWC16 logits [N,H,W,16] over the first C channels.
 * iiseg_softmax_nchw: p -> NCHW fp32 [N,C,H,W]; if y_bf16 != NULL also the
 *   NHWC bf16 [N,H,W,Cpad] copy the DAE's first conv reads.
 * iiseg_softmax_update: the loop body iterative_inference.py:267-277 for every
 *   image with active[n] != 0:  g = y - p;  y <- clip(y - step*g, 0, 1)
 *   (y NCHW fp32 master, updated in place, plus its NHWC bf16 copy);
 *   per-block partial sums of ||g||_2 over channels go to norm_partial
 *   [N][nblk] (nblk = iiseg_update_blocks(H,W)); p is also stored to p_out
 *   (NCHW fp32) when p_out != NULL.
 * iiseg_norm_finalize: norm[n] = sum(partials)/(H*W) in fixed order;
 *   n_exec[n] += 1; if norm[n] < eps: active[n] = 0 (the `break`,
 *   iterative_inference.py:275-277).  Inactive images are untouched.
 * split = 1: y_bf16 is [N,H,W,2*Cpad], the (hi | lo) bf16 pair of y (fp32-accurate variant).
 * step_dev / eps_dev != NULL: step / eps are read from device memory when the kernel runs (the by-value
 * arguments are then ignored), so a captured CUDA graph is not tied to one step size or threshold. */
int iiseg_update_blocks(int H, int W);
int iiseg_softmax_nchw(const float* logits, float* p, void* y_bf16, int N, int C,
                       int H, int W, int Cpad, int split, void* stream);
int iiseg_softmax_update(const float* logits, float* y, void* y_bf16,
                         float* p_out, const int32_t* active,
                         float* norm_partial, int N, int C, int H, int W,
                         int Cpad, float step, const float* step_dev, int split, void* stream);
/* de_fn (iterative_inference.py:203-204): grad = y - softmax(logits), NCHW fp32. */
int iiseg_softmax_grad(const float* logits, const float* y, float* grad, int N, int C,
                       int H, int W, void* stream);
int iiseg_norm_finalize(const float* norm_partial, float* norm, int32_t* active,
                        int32_t* n_exec, int N, int H, int W, float eps, const float* eps_dev,
                        void* stream);
/* Same decision from the fixed-point accumulator of the fused conv epilogue (iiseg_conv_desc.upd_*):
 * norm[n] = norm_acc[n] * 2^-40 / (H*W) for active images; norm_acc[n] is reset to 0. */
int iiseg_norm_finalize_fixed(uint64_t* norm_acc, float* norm, int32_t* active,
                              int32_t* n_exec, int N, int H, int W, float eps, const float* eps_dev,
                              void* stream);

/* ---- metrics: metrics.py jaccard / accuracy / squared_error --------------
 * One pass over y (NCHW fp32 [N,C,H,W]) and the target, per image n with
 * active[n] != 0 (active may be NULL = all), ACCUMULATING into
 *   cm     int64 [N][C*C]  cm[pred*C+true] += 1 for true < C (metrics.py:22-27)
 *   counts int64 [N][2]    {#(pred==true, true!=void), #(true!=void)} (:40-65)
 *   sqerr  fp64  [N][2]    {sum_pix mask*mean_c (y-t)^2, sum_pix mask} (:144-156)
 * argmax ties -> first index.  Target is either one-hot NCHW fp32
 * [N,C+1,H,W] (`onehot`) or int32 labels [N,H,W] (`labels`); exactly one is
 * non-NULL.  void_label < 0 means no void class. */
int iiseg_metrics_accumulate(const float* y, const float* onehot,
                             const int32_t* labels, const int32_t* active,
                             int64_t* cm, int64_t* counts, double* sqerr, int N,
                             int C, int H, int W, int void_label, void* stream);

/* labels[n,h,w] = argmax_c onehot[n,c,h,w], first index on ties: the T.argmax(y_true, axis=1)
 * of metrics.py:20-21,49-50 done once per batch.  onehot NCHW fp32 [N,C1,H,W]. */
int iiseg_onehot_to_labels(const float* onehot, int32_t* labels, int N, int C1, int H,
                           int W, void* stream);

/* ---- FC-DenseNet103 around the convs (models/FCDenseNet.py + FC_DenseNet.layers) ----------------
 * The stack of a dense block is ONE fp32 NHWC tensor [N,H,W,Cs]; "the first C channels" is the stack a
 * layer sees (ConcatLayer([stack, l]) without copies).
 * iiseg_bn_relu_pack: BatchNormLayer with batch statistics (iterative_inference.py:187,
 *   batch_norm_use_averages=False) + rectify + bf16 pack of channels [c0, c0+C):
 *   out[.., c] = bf16(relu((x - mean[c]) * (gamma[c] * inv_std[c]) + beta[c])), zero for C <= c < Cpad;
 *   mean == NULL: plain convert (TransitionUp's deconv input); relu = 0: no rectify;
 *   split = 1: out is [.., 2*Cpad], the (hi | lo) bf16 pair of the fp32 value (fp32-accurate variant).
 * iiseg_channel_stats: mean and inv_std = 1/sqrt(var + eps) (biased variance over N,H,W) of channels
 *   [c0, c0+C); deterministic two-level reduction; scratch = fp64 [iiseg_channel_stats_chunks(N,H,W)][C][2].
 * iiseg_maxpool2_f32: Pool2DLayer(2,'max') of TransitionDown on fp32 maps, written as the first C
 *   channels of the next stack (Cs_out channels per pixel).
 * iiseg_deconv_interleave: assembles Deconv2DLayer(3, stride 2, crop 'valid') from its four
 *   output-phase convolutions p[py][px] (dense fp32 [N,H+1,W+1,Cp]) with the centre crop of the
 *   following ConcatLayer: out[oh,ow] = p[(oh+crop_h)&1][(ow+crop_w)&1][(oh+crop_h)>>1, (ow+crop_w)>>1]. */
int iiseg_bn_relu_pack(const float* x, int N, int H, int W, int Cs, int c0, int C, const float* mean,
                       const float* inv_std, const float* gamma, const float* beta, int relu,
                       void* out, int Cpad, int split, void* stream);
int iiseg_channel_stats_chunks(int N, int H, int W);
int iiseg_channel_stats(const float* x, int N, int H, int W, int Cs, int c0, int C, float eps,
                        double* scratch, float* mean, float* inv_std, void* stream);
int iiseg_maxpool2_f32(const float* x, int N, int H, int W, int Cs_in, int C, float* out, int Cs_out,
                       void* stream);
int iiseg_deconv_interleave(const float* p00, const float* p01, const float* p10, const float* p11,
                            int N, int H, int W, int Cp, int C, int crop_h, int crop_w, float* out,
                            int OH, int OW, int Cs_out, void* stream);

/* ---- DAE training step (train_dae.py:243-335) around the tensor-core GEMMs -------------------------
 * iiseg_noise_pack: GaussianNoiseLayer (models/fcn_down.py:60-67) + layout change: dst = bf16 NHWC
 *   [N,H,W,Cpad] of y + sigma*noise (y, noise NCHW fp32; noise == NULL: plain pack); split = 1: [N,H,W,2*Cpad], the
 *   (hi | lo) pair.  Also used at inference for the reference's stochastic DePool2D mask sub-graph (layers/mylayers.py:91-93).
 * iiseg_loss_grad: masked crossentropy (metrics.py:68-91, clip 1e-7, void label = C) + lmb * masked
 *   squared_error (metrics.py:144-156) of softmax(logits) against the one-hot target (NCHW fp32,
 *   C+1 channels).  sums (fp64[4]) = {sum mask*CE, sum mask, sum m2*mean_c (p-t)^2, sum m2}: the loss
 *   is sums[0]/sums[1] + lmb*sums[2]/sums[3]; dlogits = bf16 NHWC16 gradient of that loss.
 *   passes: bit 0 = accumulate sums (zeroing them first), bit 1 = write dlogits using sums[1], sums[3]
 *   as the denominators (data-parallel ranks all-reduce `sums` between the two passes).
 * iiseg_loss_grad_terms: the same with the terms train_dae.py:278-294 adds up selectable: `terms` = bit 0 crossentropy |
 *   bit 1 lmb * squared_error | bit 2 dice_loss (metrics.py:93-113: -(2 I + 1) / (T + P + 1) on channel 1 of the softmax
 *   output and of the one-hot target, entries whose int32 target VALUE equals the void label id C dropped).  With bit 2
 *   `sums` is fp64[8]: sums[4] = I = sum t1 p1, sums[5] = T = sum t1, sums[6] = P = sum p1; the loss gains
 *   -(2 sums[4] + 1) / (sums[5] + sums[6] + 1).  iiseg_loss_grad == terms 3.
 * iiseg_depool2_bwd: DePool2D backward, g_u[ph,pw] = sum over the 2x2 window of mask * g_v (g_v a
 *   dense window [N,VH,VW,C] at full-resolution origin (v_h0,v_w0), zero outside; g_u dense
 *   [N,UH,UW,C] at pooled origin (u_h0,u_w0); mask full [N,H/2,W/2,C/8]).
 * iiseg_pool2_relu_bwd: Pool2DLayer(2) + rectify backward on full maps: every element that tied the
 *   window max (Theano CPU MaxPoolGrad) gets g_pool where the pooled value is > 0; where it is 0 the
 *   elements whose pre-rectifier value is exactly 0 (zmask, iiseg_conv_desc.pool_zmask) get g_pool / 2
 *   (rectify = 0.5*(x+|x|)), the negative ones 0.
 * iiseg_transpose_shift: out[(row0+c)*ldo + p] = x[n, h0+oh+dh, w0+ow+dw, c0+c] (0 outside the map),
 *   p = (n*OH+oh)*OW+ow; with nshift > 1 also the copies s = 1..nshift-1 at rows + s*shift_rows, copy s being
 *   the same matrix read at flat pixel p + s (one read of x, nshift writes): the K-major operands g^T / x^T of the weight-gradient GEMM
 *   dW[co][tap][ci] = sum_p g[p][co] x[p+tap][ci], which then runs on iiseg_conv2d_fwd (1x1, K = pixels).
 * iiseg_rmsprop_pack: lasagne.updates.rmsprop (a <- rho a + (1-rho) g^2; w <- w - lr g / sqrt(a+eps))
 *   on the fp32 master bank [Cout][taps][Cin_pad] and bias, reading g from the GEMM output [Cout][ldg]
 *   (filter tap (r, s), channel ci in column r * g_rstride + s * Cin_pad + ci, g_rstride = 0 meaning
 *   3 * Cin_pad; bias gradient in column bias_col); re-emits the bf16 forward bank wb and, if wt != NULL, the
 *   flipped / transposed bank of the data-gradient conv wt[ci-ci0][taps-1-tap][co] ([Ci_t][taps][Co_pad]). */
int iiseg_noise_pack(const float* y, const float* noise, float sigma, void* dst, int N, int C, int H,
                     int W, int Cpad, int split, void* stream);
int iiseg_loss_grad(const float* logits, const float* target, int N, int C, int H, int W, float lmb,
                    double* sums, void* dlogits, int passes, void* stream);
int iiseg_loss_grad_terms(const float* logits, const float* target, int N, int C, int H, int W,
                          float lmb, int terms, double* sums, void* dlogits, int passes, void* stream);
/* The ae_h loss term of train_dae.py:238-239,317-319: squared_error(h, h_hat).mean(), h = the DAE's own pool_{n_pool}
 * ('h_to_recon', models/fcn_down.py:117-122), h_hat = fused_up_{n_pool+1} = up_conv_{n_pool+1} + h ('h_hat', models/fcn_up.py:145-146)
 * -> the mean square of that conv's output c (the h parts cancel, in the value up to fp32 rounding and in the gradient exactly).
 * iiseg_sq_sum: sums2[0] += sum of x^2 (x bf16, n elements, n % 8 == 0), sums2[1] += n (fp64; data-parallel ranks all-reduce
 * both).  iiseg_ae_grad_add: g += 2 c / sums2[1] (g, c bf16, same n): the term's gradient with respect to c.
 * iiseg_add_bf16: out = a + b (bf16, fp32 sum, one rounding): the skip sum h_hat = c + h when c is kept on its own
 * (ElemwiseSumLayer, models/fcn_up.py:96-100; everywhere else the sum is the conv epilogue's addend). */
int iiseg_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
/* Image 0 of a batched tensor [n][image_bytes] copied to images 1..n-1 (image_bytes % 16 == 0).  The training step's
 * contracting levels above the h concat (models/fcn_down.py:77-123 with padding 100) compute the y-independent border of their
 * maps once and broadcast it; the other images run the y-dependent window only. */
int iiseg_broadcast_image(void* t, long long image_bytes, int n, void* stream);
int iiseg_sq_sum(const void* x, long long n, double* sums2, void* stream);
int iiseg_ae_grad_add(void* g, const void* c, long long n, const double* sums2, void* stream);
int iiseg_depool2_bwd(const void* gv, const uint32_t* mask, void* gu, int N, int H, int W, int C,
                      int VH, int VW, int v_h0, int v_w0, int UH, int UW, int u_h0, int u_w0,
                      void* stream);
int iiseg_pool2_relu_bwd(const void* gpool, const void* pooled, const uint32_t* mask,
                         const uint32_t* zmask, void* ga, int N, int H, int W, int C, void* stream);
int iiseg_transpose_shift(const void* x, int N, int H, int W, int Cs, int c0, int C, int h0, int w0,
                          int OH, int OW, int dh, int dw, void* out, long long ldo, long long row0,
                          int nshift, long long shift_rows, void* stream);
/* Bias gradient: out[c * ld] = sum over the P pixels of g[p][c] (bf16 NHWC rows of C channels, C % 8 == 0), fp32,
 * deterministic two-stage sum (chunk partials in `scratch`, fp32 [chunks][C], chunks <= 1024, then in chunk order). */
int iiseg_bias_grad(const void* g, long long P, int C, float* scratch, int chunks, float* out, int ld, void* stream);
/* out[i] = sum over the S slabs of in[s][i] (fp32, slab order): reduces a split-K GEMM (iiseg_conv_desc.w_koff). */
int iiseg_sum_slabs(const float* in, float* out, int S, long long n, void* stream);
int iiseg_rmsprop_pack(float* w, float* acc, float* b, float* acc_b, const float* g, void* wb, void* wt,
                       int Cout, int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0,
                       int Ci_t, int Co_pad, float lr, float rho, float eps, void* stream);
/* lasagne.updates.adam, the other optimiser of train_dae.py:326-331, on the same layouts:
 * m <- beta1 m + (1-beta1) g; v <- beta2 v + (1-beta2) g^2; w <- w - a_t m / (sqrt(v) + eps), with
 * a_t = lr sqrt(1 - beta2^t) / (1 - beta1^t) read from device memory (`a_t` = state + 1).  iiseg_adam_advance does
 * state[0] = t <- t + 1, state[1] = a_t once per step (fp32, like lasagne's shared scalar), so that a captured CUDA graph
 * replays with the right step count. */
int iiseg_adam_pack(float* w, float* m, float* v, float* b, float* m_b, float* v_b, const float* g, void* wb, void* wt,
                    int Cout, int taps, int Cin_pad, int ldg, int g_rstride, int bias_col, int ci0, int Ci_t, int Co_pad,
                    const float* a_t, float beta1, float beta2, float eps, void* stream);
int iiseg_adam_advance(float* state, float lr, float beta1, float beta2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IISEG_H_ */

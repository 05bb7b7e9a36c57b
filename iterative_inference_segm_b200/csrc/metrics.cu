// Confusion-matrix / accuracy / squared-error reduction (metrics.py) in one pass.
//
// Reference: jaccard (metrics.py:11-37) builds cm[i,j] = sum(eq(pred,i)*eq(true,j))
// with 121 separate full-array reductions; accuracy (:40-65) and squared_error
// (:144-156) are two more passes.  Here: one thread per pixel reads the C
// probability planes (coalesced NCHW), takes the first-index argmax, and the
// block accumulates a shared-memory C x C histogram with warp-aggregated
// atomics (__match_any_sync: one atomic per distinct bin per warp).  Counts are
// exact integers (int64 in HBM), so the result is order-independent and
// bit-identical across any number of GPUs.
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

constexpr int kMetBlock = 256;
constexpr int kMetMaxC = 16;

// grid = (nblk, N)
__global__ void __launch_bounds__(kMetBlock) metrics_kernel(
    const float* __restrict__ y, const float* __restrict__ onehot, const int32_t* __restrict__ labels,
    const int32_t* __restrict__ active, unsigned long long* __restrict__ cm, unsigned long long* __restrict__ counts,
    double* __restrict__ sqerr, int C, int HW, int void_label) {
  const int n = blockIdx.y;
  if (active != nullptr && active[n] == 0) return;
  __shared__ unsigned int hist[kMetMaxC * kMetMaxC];
  __shared__ unsigned int s_correct, s_valid;
  __shared__ double s_se[kMetBlock / 32], s_mask[kMetBlock / 32];
  for (int i = threadIdx.x; i < C * C; i += kMetBlock) hist[i] = 0;
  if (threadIdx.x == 0) { s_correct = 0; s_valid = 0; }
  __syncthreads();

  float se = 0.f, msum = 0.f;
  unsigned int correct = 0, valid = 0;
  for (int pix = blockIdx.x * kMetBlock + threadIdx.x; pix < HW; pix += gridDim.x * kMetBlock) {
    const float* yb = y + (size_t)n * C * HW + pix;
    float yv[kMetMaxC];
    int pred = 0; float best = yb[0]; yv[0] = best;
#pragma unroll
    for (int c = 1; c < kMetMaxC; ++c) {
      if (c < C) {
        yv[c] = yb[(size_t)c * HW];
        if (yv[c] > best) { best = yv[c]; pred = c; }   // strict >: ties keep the first index
      }
    }
    int tru;
    float pm = 0.f, pse = 0.f;   // mask = sum_c t[:, :C]; squared error over the first C channels
    if (onehot != nullptr) {
      const float* tb = onehot + (size_t)n * (C + 1) * HW + pix;
      float tbest = tb[0]; tru = 0;
#pragma unroll
      for (int c = 0; c < kMetMaxC + 1; ++c) {
        if (c <= C) {
          const float tv = tb[(size_t)c * HW];
          if (c > 0 && tv > tbest) { tbest = tv; tru = c; }
          if (c < C) { pm += tv; const float d = yv[c] - tv; pse += d * d; }
        }
      }
    } else {
      tru = labels[(size_t)n * HW + pix];
#pragma unroll
      for (int c = 0; c < kMetMaxC; ++c) {
        if (c < C) { const float tv = (c == tru) ? 1.f : 0.f; pm += tv; const float d = yv[c] - tv; pse += d * d; }
      }
    }
    se += (pse / (float)C) * pm;
    msum += pm;
    const bool is_valid = (tru != void_label);
    valid += is_valid ? 1u : 0u;
    correct += (is_valid && pred == tru) ? 1u : 0u;
    // warp-aggregated histogram update: a warp whose 32 pixels fall in one bin (the common case on real label maps)
    // issues ONE shared-memory atomic; otherwise every lane adds to its bin (match.any costs a round per distinct
    // value -- measured 43 us per call on random labels, the worst case)
    const int bin = (tru < C) ? pred * C + tru : -1;
    const unsigned int act = __activemask();
    const int leader = __ffs(act) - 1;
    const int bin0 = __shfl_sync(act, bin, leader);
    if (__all_sync(act, bin == bin0)) {
      if ((int)(threadIdx.x & 31) == leader && bin >= 0) atomicAdd(&hist[bin], (unsigned int)__popc(act));
    } else if (bin >= 0) {
      atomicAdd(&hist[bin], 1u);
    }
  }
  // block reductions
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    correct += __shfl_xor_sync(0xffffffffu, correct, o);
    valid += __shfl_xor_sync(0xffffffffu, valid, o);
  }
  double dse = (double)se, dm = (double)msum;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dse += __shfl_xor_sync(0xffffffffu, dse, o);
    dm += __shfl_xor_sync(0xffffffffu, dm, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_correct, correct); atomicAdd(&s_valid, valid);
    s_se[threadIdx.x >> 5] = dse; s_mask[threadIdx.x >> 5] = dm;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += kMetBlock)
    if (hist[i] != 0) atomicAdd(&cm[(size_t)n * C * C + i], (unsigned long long)hist[i]);
  if (threadIdx.x == 0) {
    atomicAdd(&counts[(size_t)n * 2 + 0], (unsigned long long)s_correct);
    atomicAdd(&counts[(size_t)n * 2 + 1], (unsigned long long)s_valid);
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kMetBlock / 32; ++w) { a += s_se[w]; b += s_mask[w]; }
    atomicAdd(&sqerr[(size_t)n * 2 + 0], a);
    atomicAdd(&sqerr[(size_t)n * 2 + 1], b);
  }
}

// labels[n,h,w] = argmax_c onehot[n,c,h,w] (first index on ties), as metrics.py does with one_hot=True
__global__ void __launch_bounds__(256) onehot_to_labels_kernel(const float* __restrict__ onehot, int32_t* __restrict__ labels,
                                                               int C1, int HW, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW;
    const int pix = (int)(i - n * HW);
    const float* tb = onehot + n * C1 * HW + pix;
    float best = tb[0]; int arg = 0;
    for (int c = 1; c < C1; ++c) {
      const float v = tb[(size_t)c * HW];
      if (v > best) { best = v; arg = c; }
    }
    labels[i] = arg;
  }
}

}  // namespace iiseg

extern "C" int iiseg_onehot_to_labels(const float* onehot, int32_t* labels, int N, int C1, int H, int W, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(onehot && labels, "onehot_to_labels: null tensor");
  IISEG_CHECK(N > 0 && C1 >= 1 && H > 0 && W > 0, "onehot_to_labels: bad shape");
  const long long total = (long long)N * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  onehot_to_labels_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      onehot, labels, C1, H * W, total);
  IISEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int iiseg_metrics_accumulate(const float* y, const float* onehot, const int32_t* labels,
                                        const int32_t* active, int64_t* cm, int64_t* counts, double* sqerr, int N,
                                        int C, int H, int W, int void_label, void* stream) {
  using namespace iiseg;
  IISEG_CHECK(y && cm && counts && sqerr, "metrics: null tensor");
  IISEG_CHECK((onehot != nullptr) != (labels != nullptr), "metrics: exactly one of onehot/labels must be given");
  IISEG_CHECK(N > 0 && C >= 1 && C <= kMetMaxC && H > 0 && W > 0, "metrics: bad shape");
  const int HW = H * W;
  int nblk = (HW + kMetBlock - 1) / kMetBlock;
  int cap = num_sms() * 4 / N;                            // one resident wave across the batch (measured: 4 and 8 blocks per SM tie, 2 is 1.7x slower)
  if (cap < 1) cap = 1;
  if (nblk > cap) nblk = cap;
  dim3 grid(nblk, N);
  metrics_kernel<<<grid, kMetBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y, onehot, labels, active, reinterpret_cast<unsigned long long*>(cm), reinterpret_cast<unsigned long long*>(counts),
      sqerr, C, HW, void_label);
  IISEG_LAUNCH_CHECK();
  return 0;
}

// Experiment (building block of "N-packing"): does a TMA tensor-map box with elementStrides = 4 along
// the W axis of an NHWC bf16 tensor deliver every fourth pixel as consecutive 128-byte SWIZZLE_128B rows?
// Tensor [N=1][H=8][W=64][C=64]; box = 64 channels x (16 pixels spanning 64 with stride 4) x 2 rows, loaded at
// W coordinate c0 = 0..3.  Expected smem row r = (h, m): pixel (h, 4m + c0), chunk q at ((q ^ (r & 7)) << 4).
// RESULT (B200): yes.  boxDim[W] must be the number of elements SPANNED (16 pixels x stride 4 = 64), elementStrides[W] = 4;
// every start coordinate c0 = 0..3 works (W is not the contiguous axis) and the 128B swizzle is the usual address-based one.
// With boxDim[W] = 16 only 4 pixels per line arrive (the transaction-byte count then differs: a bounded wait times out).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_test tma_stride_test.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* out, int c0, int rows) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(base)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(0), "r"(c0), "r"(1), "r"(0) : "memory");
  }
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT\n"); break; }
  }
  for (int i = threadIdx.x; i < rows * 64; i += 128) {           // un-swizzle into out[row][ch]
    const int r = i / 64, ch = i % 64;
    out[i] = *reinterpret_cast<__nv_bfloat16*>(base + r * 128 + ((((ch >> 3) ^ (r & 7))) << 4) + (ch & 7) * 2);
  }
}

int main() {
  const int H = 8, W = 64, C = 64, MB = 16, TH = 2;
  std::vector<__nv_bfloat16> h(H * W * C);
  for (int i = 0; i < H * W * C; ++i) h[i] = __float2bfloat16((float)((i / C) % 251) + 0.001f * 0);   // value = pixel index mod 251 (exact in bf16 up to 256)
  __nv_bfloat16 *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, TH * MB * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  int bad_total = 0;
  for (int variant = 0; variant < 2; ++variant) {
    // variant 0: boxDim[W] = MB * 4 (elements spanned), variant 1: boxDim[W] = MB (elements delivered)
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(variant == 0 ? MB * 4 : MB), (cuuint32_t)TH, 1};
    cuuint32_t estr[4] = {1, 4, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d (boxDim[W] = %u): encode -> %d\n", variant, box[1], (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int c0 = 0; c0 < 4; ++c0) {
      cudaMemset(o, 0xFF, TH * MB * 64 * 2);
      k<<<1, 128, 48 * 1024>>>(tm, o, c0, TH * MB);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  c0=%d: launch failed: %s\n", c0, cudaGetErrorString(e)); return 1; }
      std::vector<__nv_bfloat16> res(TH * MB * 64);
      cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < TH * MB; ++r2) {
        const int hh = 1 + r2 / MB, m = r2 % MB, pix = hh * W + 4 * m + c0;
        const float want = (float)(pix % 251);
        for (int ch = 0; ch < 64; ++ch) if (__bfloat162float(res[r2 * 64 + ch]) != want) { if (bad < 3) printf("  c0=%d row %d ch %d: got %g want %g\n", c0, r2, ch, __bfloat162float(res[r2 * 64 + ch]), want); ++bad; }
      }
      printf("  c0=%d: %s (%d mismatches)\n", c0, bad ? "MISMATCH" : "rows = every 4th pixel, as expected", bad);
      bad_total += bad;
    }
  }
  return bad_total ? 2 : 0;
}

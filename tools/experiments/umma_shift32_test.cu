// Experiment: K-major SWIZZLE_32B operands (16 bf16 = 32 B per row, one K=16 MMA per row block): does
// tcgen05.mma accept an A descriptor whose start address is shifted by s rows (s*32 B) inside the 256-byte
// swizzle atom?  A is written with the address-based 32B swizzle TMA uses (16-byte chunk ^= address bit 7).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int use_base_offset, int a_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  uint8_t* sA = base;                 // a_rows x 128 B (a_rows <= 144)
  uint8_t* sB = base + 20480;         // 64 x 128 B
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  // fill A, B with swizzle: element (row, k) -> row*128 + ((k/8 ^ (row&7))*16) + (k%8)*2
  for (int i = tid; i < a_rows * 16; i += 128) {
    int row = i / 16, kk = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(sA + row * 32 + (((kk >> 3) ^ ((row >> 2) & 1)) << 4) + (kk & 7) * 2) = A[i];
  }
  for (int i = tid; i < 64 * 16; i += 128) {
    int row = i / 16, kk = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(sB + row * 32 + (((kk >> 3) ^ ((row >> 2) & 1)) << 4) + (kk & 7) * 2) = B[i];
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(sA) + shift * 32, b_addr = smem_u32(sB);
    uint64_t a_desc = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
    if (use_base_offset) a_desc |= (uint64_t)((a_addr >> 7) & 1) << 49;
    uint64_t b_desc = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    {
      uint32_t acc = 0;
      asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                   ::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok) {
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (tid == 0) printf("TIMEOUT\n"); break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < 64; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
  const int AR = 176;
  __nv_bfloat16 *hA = new __nv_bfloat16[AR * 16], *hB = new __nv_bfloat16[64 * 16];
  float* fA = new float[AR * 16]; float* fB = new float[64 * 16];
  srand(1);
  for (int i = 0; i < AR * 16; ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 64 * 16; ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, AR * 16 * 2); cudaMalloc(&dB, 64 * 16 * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA, AR * 16 * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 64 * 16 * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  float* hD = new float[128 * 64];
  for (int bo = 0; bo < 2; ++bo)
    for (int si = 0; si < 16; ++si) {
      const int shifts[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 18, 32, 33, 34};
      const int shift = shifts[si];
      cudaMemset(dD, 0, 128 * 64 * 4);
      k<<<1, 128, 64 * 1024>>>(dA, dB, dD, shift, bo, AR);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0, bad_rowmod[8] = {0};
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int kk = 0; kk < 16; ++kk) ref += fA[(m + shift) * 16 + kk] * fB[n * 16 + kk];
          if (fabsf(ref - hD[m * 64 + n]) > 1e-3f) { ++bad; bad_rowmod[m & 7]++; }
        }
      printf("shift %2d base_offset_field %d: %s (%d mismatches; by m%%8: %d %d %d %d %d %d %d %d)\n", shift, bo, bad ? "WRONG" : "OK", bad,
             bad_rowmod[0], bad_rowmod[1], bad_rowmod[2], bad_rowmod[3], bad_rowmod[4], bad_rowmod[5], bad_rowmod[6], bad_rowmod[7]);
    }
  return 0;
}

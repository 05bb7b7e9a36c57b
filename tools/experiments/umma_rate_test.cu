// Experiment: tcgen05.mma issue-to-retire rate (cycles per MMA) for M=128, K=16, N in {16,64,128,256},
// with the A descriptor start aligned (shift 0) or advanced by 1/2/8 rows inside the SWIZZLE_128B atom.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int shift, int reps) {
  extern __shared__ __align__(1024) uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 60 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw)[i] = 0x3c003c00u;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(raw) + shift * 128, b_addr = smem_u32(raw) + 24 * 1024;
    const uint64_t a_desc = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint64_t b_desc = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      uint32_t acc = 1;
      asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                   ::"r"(tmem), "l"(a_desc + 2 * (i & 3)), "l"(b_desc + 2 * (i & 3)), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2048;
  int Ns[4] = {16, 64, 128, 256}, shifts[8] = {0, 1, 2, 8, 64, 65, 66, 130};
  for (int ni = 0; ni < 4; ++ni)
    for (int si = 0; si < 8; ++si) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 64 * 1024>>>(d, Ns[ni], shifts[si], reps); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d shift=%d: %.1f cycles / MMA (128xNx16)\n", Ns[ni], shifts[si], (double)h / reps);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

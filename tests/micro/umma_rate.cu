// Hardware experiment (test infrastructure): sustained tcgen05.mma rate (cycles per M=128 x N x K=16 bf16 MMA) as a
// function of N, of the A descriptor form (aligned 1024-byte atoms vs row-shifted halo windows with SBO = 1280) and of
// the number of CTAs per GPU.  Explains the per-MMA cost seen in the halo kernel timeline (profiles/).
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// mode 0: aligned A tile (SBO 1024), same descriptor set each tap
// mode 1: halo windows: tap (ky,kx) start = (ky*10+kx)*128, SBO 1280
// mode 2: like 1 but every MMA uses a different accumulator column offset (independent accumulators, round robin of 2)
__global__ void __launch_bounds__(128, 1) k_rate(int n, int mode, int rounds, int split3, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  // [A hi: 23 KB][A lo: 23 KB][B hi: n*128][B lo: n*128]
  const uint32_t a_hi = base, a_lo = base + 23 * 1024, b_hi = base + 46 * 1024, b_lo = b_hi + 256 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 46 * 1024 + 2 * 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < (46 * 1024 + 2 * 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i % 7;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sbo = mode == 0 ? 1024 : 1280;
    long long t0 = clock64();
    int cnt = 0;
    for (int r = 0; r < rounds; ++r) {
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t aoff = mode == 0 ? 0 : (uint32_t)((tap / 3) * 10 + tap % 3) * 128u;
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = desc(a_hi + aoff + k * 32, sbo), dal = desc(a_lo + aoff + k * 32, sbo);
          const uint64_t dbh = desc(b_hi + k * 32, 1024), dbl = desc(b_lo + k * 32, 1024);
          const uint32_t tacc = tm + (mode == 2 ? (uint32_t)((cnt & 1) * n) : 0u);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc),
                       "l"(dal), "l"(dbh), "r"(idesc), "r"(1)
                       : "memory");
          ++cnt;
          if (split3) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc),
                         "l"(dah), "l"(dbl), "r"(idesc), "r"(1)
                         : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc),
                         "l"(dah), "l"(dbh), "r"(idesc), "r"(1)
                         : "memory");
            cnt += 2;
          }
        }
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
      cycles[2] = cnt;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 3 * sizeof(long long));
  long long h[3];
  const int smem = 46 * 1024 + 2 * 256 * 128 + 64 + 1024;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int ns[] = {32, 64, 128, 256};
  for (int grid : {1, 148})
    for (int split3 : {0, 1})
      for (int mode = 0; mode < 3; ++mode)
        for (int n : ns) {
          if (mode == 2 && 2 * n > 512) continue;
          k_rate<<<grid, 128, smem>>>(n, mode, 20, split3, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          printf("grid %3d split3 %d mode %d N %3d : %lld MMAs, issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor N/2 = %d)\n", grid, split3, mode, n,
                 h[2], (double)h[0] / h[2], (double)h[1] / h[2], n / 2);
        }
  return 0;
}

// Hardware experiment (test infrastructure): does a K-major SWIZZLE_128B UMMA operand descriptor accept a start
// address that is a multiple of 128 B but NOT of 1024 B (a window shifted by whole rows inside a larger swizzled
// buffer), and with which `base_offset` / SBO?  Needed for the halo-tile convolution (9 taps = 9 shifted windows).
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o umma_shift umma_shift.cu && ./umma_shift
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_shift(int shift_rows, int sbo_bytes, int base_off, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* A = sm;                    // 512 rows x 128 B
  uint8_t* B = sm + 512 * 128;        // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < 512 * 64; i += 128) {
    int hp = i / 64, k = i % 64;
    float v = (float)(((hp * 7 + k * 3) % 17) - 8);
    *reinterpret_cast<__nv_bfloat16*>(A + hp * 128 + (((k / 8) ^ (hp & 7)) << 4) + (k % 8) * 2) = __float2bfloat16(v);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    int n = i / 64, k = i % 64;
    float v = (float)(((n * 5 + k) % 13) - 6);
    *reinterpret_cast<__nv_bfloat16*>(B + n * 128 + (((k / 8) ^ (n & 7)) << 4) + (k % 8) * 2) = __float2bfloat16(v);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < 4; ++j) {
      uint32_t a_addr = smem_u32(A) + shift_rows * 128 + j * 32, b_addr = smem_u32(B) + j * 32;
      uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
                    ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
      uint64_t db = (uint64_t)((b_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      uint32_t acc = j != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                   "l"(da), "l"(db), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait for the MMAs
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  float* h = (float*)malloc(128 * 64 * 4);
  const int smem = 512 * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int shifts[] = {0, 1, 3, 8, 11, 21};
  int sbos[] = {1024, 1280, 2304};
  for (int sbo : sbos)
    for (int sh : shifts)
      for (int mode = 0; mode < 2; ++mode) {
        int bo = mode ? (sh & 7) : 0;
        if (mode && bo == 0) continue;
        cudaMemset(d, 0, 128 * 64 * 4);
        k_shift<<<1, 128, smem>>>(sh, sbo, bo, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d base_off %d: CUDA error %s\n", sbo, sh, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        // expected: MMA row m <-> halo row shift + (m/8)*(sbo/128) + m%8
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m) {
          int hp = sh + (m / 8) * (sbo / 128) + (m % 8);
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)(((hp * 7 + k * 3) % 17) - 8) * (double)(((n * 5 + k) % 13) - 6);
            double err = fabs(ref - h[m * 64 + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) ++bad;
          }
        }
        printf("sbo %4d shift %2d base_offset %d : max err %.3g, wrong %d / 8192 -> %s\n", sbo, sh, bo, maxerr, bad, bad ? "MISMATCH" : "OK");
      }
  return 0;
}

// PTX wrappers and tile constants shared by the tcgen05 convolution kernels (conv_tc.cu, conv_tc_halo.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace e4s {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 2) * 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // one bf16 A tile (hi or lo): 16 KB

__host__ __device__ constexpr int tc_block_n(int cout) { return cout >= 256 ? 256 : cout; }
__host__ __device__ constexpr int tc_stage_bytes(int bn) { return 2 * TC_A_BYTES + 2 * bn * TC_BK * 2; }
__host__ __device__ constexpr int tc_stages(int bn) { return bn == 256 ? 2 : (bn == 128 ? 3 : (bn == 64 ? 4 : 5)); }
__host__ __device__ constexpr int tc_smem_bytes(int bn) { return tc_stages(bn) * tc_stage_bytes(bn) + TC_BM * 16 + 256 + 3 * bn * 4 + 1024; }   // + [mul | add | slope] epilogue vectors

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait suspends in hardware for a bounded time per call; the spin bound turns a protocol bug
// (a barrier that never completes) into a trap -> launch error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("e4s conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// ---- thread-block clusters: weight tiles are fetched from L2 once per cluster and multicast into every CTA's smem ------
__device__ __forceinline__ uint32_t cluster_nctas() {
  uint32_t n;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
  return n;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> the same shared-memory offset of every CTA in cta_mask; complete_tx lands on the mbarrier at the same offset in each
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
               : "memory");
}
// arrive (once the MMAs issued so far have completed) on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}

// ---- CTA pairs (cta_group::2): two SMs of a TPC execute one M=256 MMA, each holding its own A rows and HALF of B ----------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta_rank) {     // my shared address -> the same offset in cta_rank
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}
// wait on a barrier whose arrivals come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 22)) {
      printf("e4s conv_tc: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive accumulator columns of this thread's TMEM lane, no wait (pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// the registers are threaded through the wait as in/out operands so that no use can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld16x2(uint32_t (&r)[16], uint32_t (&q)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(q[0]), "+r"(q[1]), "+r"(q[2]),
                 "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9]), "+r"(q[10]), "+r"(q[11]),
                 "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])
               :
               : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  Field layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: c=f32 (bit 4), a=bf16 (bit 7), b=bf16 (bit 10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// one elected lane of a fully converged warp (PTX elect.sync); keeps the surrounding control flow warp-uniform so the
// compiler emits single UTCHMMA / UBLKCP instructions instead of a per-active-lane issue loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// streaming activation load: read-only path, do not allocate in L1 (the tiny L1 left beside ~200 KB of shared memory
// must keep the per-channel epilogue vectors and style rows resident)
__device__ __forceinline__ float4 ldg_stream4(const float4* ptr) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr));
  return r;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);  // .x (low 16 bits) = lo_elem
  return *reinterpret_cast<uint32_t*>(&v);
}

// The 3-pass operand split v = hi + lo of two consecutive K elements (a = lower address), as two packed 16-bit pairs.
//   bf16 (E4S_TC_BF16): 8 + 8 mantissa bits, fp32 exponent range  -> 2^-17 relative per operand
//   fp16 (E4S_TC_F16):  11 + 11 mantissa bits, |v| <= 65504 (clamped; lo parts below 2^-14 fall into the fp16 subnormals,
//                       absolute resolution 2^-25)                    -> 2^-22 relative per operand: fp32 class
// Both run as tcgen05.mma kind::f16 at the same rate; only the a/b format bits of the instruction descriptor differ.
//
// fp16 mode also SEPARATES the accumulators (TC_LO_SCALE): the tensor core truncates every product to the accumulator's ulp when
// it aligns the addends, so adding the 2^-11-sized hi*lo / lo*hi terms into the main accumulator costs two more ulp-sized
// truncations per element (measured: 3.4x the error of an fp32 FMA chain at K = 4608 although the operands are exact to 2^-22).
// Here hi*hi accumulates alone and the two small terms go to a second TMEM accumulator whose truncation unit is 2^-11 of the
// main one's; the lo parts are stored multiplied by 2^11 (never subnormal when hi is normal) and the epilogue adds
// small * 2^-11 to the main sum in fp32.
constexpr float TC_LO_SCALE = 2048.f;
__device__ __forceinline__ void tc_split2(const bool f16, const float a, const float b, uint32_t& hi, uint32_t& lo) {
  if (f16) {
    const float ac = fminf(fmaxf(a, -65504.f), 65504.f), bc = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(ac, bc);                 // .x (low 16 bits) = a
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn((ac - back.x) * TC_LO_SCALE, (bc - back.y) * TC_LO_SCALE);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const uint32_t h = pack_bf16x2(a, b);
    hi = h;
    lo = pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xffff0000u));
  }
}
// instruction-descriptor a/b format fields: 1 = bf16 (what umma_idesc sets), 0 = fp16
__host__ __device__ __forceinline__ uint32_t umma_idesc_fmt(uint32_t idesc_bf16, int tc_fmt) {
  return tc_fmt == E4S_TC_F16 ? idesc_bf16 & ~((1u << 7) | (1u << 10)) : idesc_bf16;
}

// ---------------------------------------------------------------------------------------------------
// fused epilogue for one accumulator row (one output pixel) x 16 consecutive channels, shared by the tcgen05
// kernels.  Per-channel vectors are read with 16-byte read-only loads (L1 resident), the pixel's 16 outputs leave
// as four st.global.v4.  Kept compact on purpose: the first version unrolled ~40 scalar instructions per element
// and the epilogue warps became the bottleneck of the small-N layers (profiles/r1_ncu_halo_v1_stalls.txt).
// ---------------------------------------------------------------------------------------------------
struct TcEpiRow {
  const float* drow;     // demodulation row (b, region) or nullptr
  const float* nrow;     // noise plane pointer at this pixel (channel 0) or nullptr
  float pw;              // float-mask weight (generic path) when p.pixw
  float nz;              // noise_w * noise[pixel] when the noise has one channel
  float nw;              // noise weight
  int64_t pix;           // output pixel index
};

__device__ __forceinline__ float4 ldg4(const float* ptr) { return __ldg(reinterpret_cast<const float4*>(ptr)); }

// NOTE: no `switch` here -- nvcc lowers it to a jump table (LDC + BRX indirect branch) per element, which made the
// epilogue the bottleneck of the small-N layers (2.2 of 5.8 ms on the 32->32 @1024^2 layer).
__device__ __forceinline__ float4 tc_act4(float4 a, const int act, const float slope, const float gain) {
  if (act <= E4S_ACT_RELU) {
    // NONE / LRELU / RELU as one branch-free form: (t < 0 ? t*sl : t) * g with (sl, g) = (1,1) / (slope, gain) / (1,1) and a lower clamp at 0 for RELU
    const float sl = act == E4S_ACT_LRELU ? slope : 1.f;
    const float g = act == E4S_ACT_LRELU ? gain : 1.f;
    const float lo = act == E4S_ACT_RELU ? 0.f : -INFINITY;          // RELU clamps at 0 (select, no branch)
    a.x = fmaxf((a.x < 0.f ? a.x * sl : a.x) * g, lo); a.y = fmaxf((a.y < 0.f ? a.y * sl : a.y) * g, lo);
    a.z = fmaxf((a.z < 0.f ? a.z * sl : a.z) * g, lo); a.w = fmaxf((a.w < 0.f ? a.w * sl : a.w) * g, lo);
  } else if (act == E4S_ACT_SIGMOID) {
    a.x = 1.f / (1.f + expf(-a.x)); a.y = 1.f / (1.f + expf(-a.y)); a.z = 1.f / (1.f + expf(-a.z)); a.w = 1.f / (1.f + expf(-a.w));
  } else {   // E4S_ACT_RSQRT_EPS (PReLU is handled by the caller)
    a.x = rsqrtf(a.x + slope); a.y = rsqrtf(a.y + slope); a.z = rsqrtf(a.z + slope); a.w = rsqrtf(a.w + slope);
  }
  return a;
}

// One float4 group of the epilogue given the (already fetched) per-channel vectors.
__device__ __forceinline__ float4 tc_epi_math4(const E4SConv& p, float4 a, const TcEpiRow& r, const int n, const float4 mul, const float4 add,
                                               const float4 prelu) {
  a.x *= mul.x; a.y *= mul.y; a.z *= mul.z; a.w *= mul.w;          // demod * channel scale
  if (p.pixw) {
    a.x *= r.pw; a.y *= r.pw; a.z *= r.pw; a.w *= r.pw;
  }
  if (r.nrow) {
    if (p.noise_sc == 0) {
      a.x += r.nz; a.y += r.nz; a.z += r.nz; a.w += r.nz;
    } else {
      a.x += r.nw * __ldg(r.nrow + (int64_t)n * p.noise_sc);
      a.y += r.nw * __ldg(r.nrow + (int64_t)(n + 1) * p.noise_sc);
      a.z += r.nw * __ldg(r.nrow + (int64_t)(n + 2) * p.noise_sc);
      a.w += r.nw * __ldg(r.nrow + (int64_t)(n + 3) * p.noise_sc);
    }
  }
  a.x += add.x; a.y += add.y; a.z += add.z; a.w += add.w;          // bias / BN shift
  float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.res) {
    rs = ldg4(p.res + r.pix * p.res_pitch + n);
    if (!p.res_after_act) {
      a.x += rs.x; a.y += rs.y; a.z += rs.z; a.w += rs.w;
    }
  }
  if (p.act == E4S_ACT_PRELU) {
    a.x = a.x < 0.f ? a.x * prelu.x : a.x; a.y = a.y < 0.f ? a.y * prelu.y : a.y;
    a.z = a.z < 0.f ? a.z * prelu.z : a.z; a.w = a.w < 0.f ? a.w * prelu.w : a.w;
  } else {
    a = tc_act4(a, p.act, p.act_slope, p.act_gain);
  }
  if (p.res && p.res_after_act) {
    a.x += rs.x; a.y += rs.y; a.z += rs.z; a.w += rs.w;
  }
  return a;
}

__device__ __forceinline__ void tc_epi_store4(const E4SConv& p, float4* o, float4 a) {
  if (p.accumulate) {
    const float4 old = *o;
    a.x += old.x; a.y += old.y; a.z += old.z; a.w += old.w;
  }
  *o = a;
}

// per-channel vectors from global memory (per-row demodulation: gather kernel).  All loads are issued before the
// first use so their L2 latency overlaps (they were serialised per float4 group in the first version).
__device__ __forceinline__ void tc_epilogue16(const E4SConv& p, const float (&acc)[16], const int n0, const TcEpiRow& r) {
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 mul[4], add[4], pre[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + 4 * q;
    mul[q] = r.drow ? ldg4(r.drow + n) : one;
    if (p.ch_scale) {
      const float4 s = ldg4(p.ch_scale + n);
      mul[q].x *= s.x; mul[q].y *= s.y; mul[q].z *= s.z; mul[q].w *= s.w;
    }
    add[q] = p.ch_shift ? ldg4(p.ch_shift + n) : zero;
    pre[q] = p.act == E4S_ACT_PRELU ? ldg4(p.act_prelu + n) : zero;
  }
  float4* o = reinterpret_cast<float4*>(p.out + r.pix * p.out_pitch + n0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    tc_epi_store4(p, o + q, tc_epi_math4(p, a, r, n0 + 4 * q, mul[q], add[q], pre[q]));
  }
}

// per-channel vectors staged in shared memory by the caller: sv = [mul(BN) | add(BN) | prelu(BN)], nl = local channel
// dbg (profiling experiments): bit3 = skip the global stores, bit4 = skip the math (store the raw accumulator)
__device__ __forceinline__ void tc_epilogue16_sv(const E4SConv& p, const float (&acc)[16], const int n0, const int nl, const int bn,
                                                 const float* sv, const TcEpiRow& r, const int dbg = 0) {
  float4* o = reinterpret_cast<float4*>(p.out + r.pix * p.out_pitch + n0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 a = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    if (!(dbg & 16)) {
      const float4 mul = *reinterpret_cast<const float4*>(sv + nl + 4 * q);
      const float4 add = *reinterpret_cast<const float4*>(sv + bn + nl + 4 * q);
      const float4 pre = *reinterpret_cast<const float4*>(sv + 2 * bn + nl + 4 * q);
      a = tc_epi_math4(p, a, r, n0 + 4 * q, mul, add, pre);
    }
    if (!(dbg & 8) || a.x == 123456.789f) tc_epi_store4(p, o + q, a);
  }
}

// The tensor core's fp32 accumulate truncates: every tcgen05.mma into the same TMEM accumulator shrinks the running sum by
// ~2^-26 relative on average (measured, tests/micro/acc_bias.py: least-squares scale of the result against fp64 is
// -1.5e-8 x accumulate steps for K = 576 ... 4608, while the CUDA-core fp32 engine shows none).  A K = 4608 layer loses
// 1.3e-5, and because the synthesis network is positively homogeneous the per-layer losses ADD UP along the 17 layers
// into the dominant, coherent part of the end-to-end error.  The epilogues multiply the accumulator by this factor to
// remove the mean of that bias (the random part of the truncation stays).
// fp16 operands (22-bit products) leave bits below the accumulator's ulp at almost every step, bf16 products (16 bits) only when they
// are much smaller than the running sum: the per-step constant is per format (E4SConv.tc_unbias overrides it; < 0 disables).
constexpr float TC_UNBIAS_BF16 = 1.5e-8f;
constexpr float TC_UNBIAS_F16 = 1.7e-8f;       // 1.6e-8 (zero-mean operands) ... 2.5e-8 (same-signed sums), profiles/r2_acc_bias.txt
__host__ __device__ __forceinline__ float tc_acc_unbias(const E4SConv& p, int mma_steps) {
  const float per_step = p.tc_unbias != 0.f ? fmaxf(p.tc_unbias, 0.f) : (p.tc_fmt == E4S_TC_F16 ? TC_UNBIAS_F16 : TC_UNBIAS_BF16);
  return (1.f + per_step * (float)mma_steps) * (p.tc_out_scale != 0.f ? p.tc_out_scale : 1.f);
}

// Fast epilogue (chosen once per kernel, outside every loop): no residual / float mask / accumulate / per-channel noise and an
// activation of the piecewise-linear family.  sv = [mul | add | negative-side slope] per channel; NONE, ReLU, leaky ReLU
// and PReLU all are  t = acc*mul + (add + nz);  out = (t < 0 ? t*slope : t) * gain  -- no branches, no constant-bank
// loads per element (the generic path spends ~70 instructions per float4 on uniform branches).
__host__ __device__ __forceinline__ bool tc_epi_is_fast(const E4SConv& p) {
  return !p.res && !p.pixw && !p.accumulate && (!p.noise || p.noise_sc == 0) &&
         (p.act == E4S_ACT_NONE || p.act == E4S_ACT_RELU || p.act == E4S_ACT_LRELU || p.act == E4S_ACT_PRELU);
}
__device__ __forceinline__ float tc_epi_slope(const E4SConv& p, const int n) {
  return p.act == E4S_ACT_PRELU ? __ldg(p.act_prelu + n) : (p.act == E4S_ACT_LRELU ? p.act_slope : (p.act == E4S_ACT_RELU ? 0.f : 1.f));
}
template <int NV, bool RGB = false>   // NV consecutive channels (multiple of 4) of one pixel
__device__ __forceinline__ void tc_epilogue_fast(float* __restrict__ optr, const float (&acc)[NV], const float* sv, const int nl, const int bn,
                                                 const float nz, const float gain, const float* rgbw = nullptr, float* rgb3 = nullptr) {
  float4* o = reinterpret_cast<float4*>(optr);
#pragma unroll
  for (int q = 0; q < NV / 4; ++q) {
    const float4 mul = *reinterpret_cast<const float4*>(sv + nl + 4 * q);
    const float4 add = *reinterpret_cast<const float4*>(sv + bn + nl + 4 * q);
    const float4 sl = *reinterpret_cast<const float4*>(sv + 2 * bn + nl + 4 * q);
    float4 a;
    a.x = fmaf(acc[4 * q], mul.x, add.x + nz); a.y = fmaf(acc[4 * q + 1], mul.y, add.y + nz);
    a.z = fmaf(acc[4 * q + 2], mul.z, add.z + nz); a.w = fmaf(acc[4 * q + 3], mul.w, add.w + nz);
    a.x = (a.x < 0.f ? a.x * sl.x : a.x) * gain; a.y = (a.y < 0.f ? a.y * sl.y : a.y) * gain;
    a.z = (a.z < 0.f ? a.z * sl.z : a.z) * gain; a.w = (a.w < 0.f ? a.w * sl.w : a.w) * gain;
    if (RGB) {
      // fused ToRGB: three running dot products with the sample's modulated 1x1 weights (rgbw = [3][bn] in shared memory)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(rgbw + c * bn + nl + 4 * q);
        rgb3[c] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, rgb3[c]))));
      }
    }
    if (optr) o[q] = a;
  }
}

}  // namespace e4s

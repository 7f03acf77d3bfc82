// tcgen05 "halo" convolution: 3x3 stride-1 convolutions (and the poly-phase up-convolution) whose A-operand
// transform is uniform over a tile -- the un-masked StyleGAN2 layers (one style per sample: 512^2 / 1024^2, the
// HBM-heavy layers) and the encoder's stride-1 convs (InstanceNorm per sample/channel).
//
// The gather engine (conv_tc.cu) rebuilds an im2col A tile per tap: 9x (36x for up-convs) redundant loads,
// normalise/modulate/split work and shared-memory stores -- it is LSU/latency bound on the small-N layers
// (profiles/r1_ncu_conv_tc_v1_per_layer.csv: 3-8 % tensor pipe at 512^2/1024^2).  Here the producers convert the
// (16+2) x (8+2) pixel input HALO of a 16x8 output tile ONCE per 64-channel group into bf16 hi/lo planes stored
// pixel-major with the 128-byte swizzle on absolute shared-memory addresses, and the 9 taps are 9 UMMA descriptors
// whose start address is shifted by (ky*10 + kx) pixel rows with SBO = one halo row (1280 B).  tests/micro/
// umma_shift.cu established on hardware that the tensor core applies the swizzle to absolute address bits, so
// such row-shifted windows (start not 1024-aligned, base_offset 0, arbitrary SBO) read back exactly.
// The four phases of an up-convolution reuse the same halo with four accumulators in TMEM.
//
// Persistent CTAs (one per SM), 448 threads:
//   warps 0-7   halo producers: coalesced NHWC fp32 loads -> InstanceNorm / style modulation -> bf16 hi/lo split ->
//               swizzled st.shared; one-job-ahead register prefetch
//   warps 8-11  epilogue: tcgen05.ld -> demod / noise / bias / residual / activation -> NHWC fp32 stores
//   warp 12     MMA issuer + TMEM owner (accumulator sets double-buffered when 2*P*BN <= 512 columns)
//   warp 13     weight loader: cp.async.bulk of the pre-swizzled bf16 hi|lo tiles into a ring of stages
#include "tc_ptx.cuh"

namespace e4s {

constexpr int HL_TH = 16, HL_TW = 8;
constexpr int HL_HP = HL_TW + 2;                  // halo row pitch (pixels)
constexpr int HL_HPIX = (HL_TH + 2) * HL_HP;      // 180 halo pixels
constexpr int HL_PLANE = HL_HPIX * 128;            // one bf16 plane (hi or lo): 180 pixel rows of 128 B (NOT 1 KB aligned:
                                                  // the swizzle phase is taken from the absolute shared-memory address)
constexpr int HL_HALO_STAGES = 2;
constexpr int HL_THREADS = 14 * 32;
constexpr int HL_MMA_WARP = 12;             // warps 8-11 epilogue, 12 MMA issuer, 13 weight loader
constexpr int HL_ITEMS = (HL_HPIX + 31) / 32;     // 6 (pixel, 8-channel) items per producer thread

constexpr int HL_MAX_BST = 6;
constexpr int HL_SMEM_MAX = 232448;               // 227 KB opt-in limit per CTA
// phases of an up-convolution merged into the N dimension of one MMA (all four phases read the SAME halo window)
__host__ __device__ constexpr int hl_phase_merge(int bn, bool up) { return (up && 4 * bn <= 256) ? 4 : 1; }
// fixed part: halo stages + 2 slots of 6 per-channel epilogue vectors (mul | add | slope | 3 ToRGB rows) + barriers + alignment slack
__host__ __device__ constexpr int hl_fixed_bytes(int bn) { return HL_HALO_STAGES * 2 * HL_PLANE + 2 * 6 * bn * 4 + 512 + 1024; }
__host__ __device__ constexpr int hl_b_stages(int bn, int pm) {
  int n = (HL_SMEM_MAX - hl_fixed_bytes(bn)) / (2 * pm * bn * 128);
  return n > HL_MAX_BST ? HL_MAX_BST : n;
}
__host__ __device__ constexpr int hl_smem_bytes(int bn, int pm) { return hl_fixed_bytes(bn) + hl_b_stages(bn, pm) * 2 * pm * bn * 128; }

__device__ __forceinline__ uint64_t umma_smem_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct HlJob {
  int b, y0, x0, nt, reg;
};

// optional cycle accounting of CTA 0 (profiling aid): one lead thread per role adds the clock64 time between consecutive
// laps to a shared-memory counter [role][lap]; the totals are written to the installed buffer when the kernel ends.  Unlike
// a global-memory event trace this costs a few cycles per lap, so the numbers are those of the undisturbed kernel.
__device__ unsigned long long* g_hl_trace = nullptr;
__device__ int g_hl_trace_cap = 0;
__device__ int g_hl_dbg = 0;      // profiling experiments: bit0 skip epilogue math+stores, bit1 skip tcgen05.ld, bit2 skip MMAs
#ifdef E4S_HL_ACCT                  // E4S_HL_ACCT=1 python -m e4s2024_b200.build --force  (tests/micro/halo_trace.py)
#define HL_LAP(role, k)                                   \
  do {                                                    \
    if (acct) {                                           \
      const long long n_ = clock64();                     \
      s_acct[(role) * 8 + (k)] += n_ - acct_t;            \
      acct_t = n_;                                        \
    }                                                     \
  } while (0)
#else                               // the predicated-off laps still cost ~20 issue slots per weight chunk in the MMA warp
#define HL_LAP(role, k) \
  do {                  \
  } while (0)
#endif

// UMMA descriptor words.  The 64-bit shared-memory descriptor is affine in the byte address through its low word only
// (14-bit start-address field in 16-byte units, smem < 256 KB => no carry), so the issue loop does 32-bit adds on the low
// word and the high word (SBO, version, swizzle mode) is a constant.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_bf16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN, bool UP, bool C32>
__global__ void __launch_bounds__(HL_THREADS, 1)   // 14 warps are allocated as 16: 128 registers per thread is the ceiling
conv_tc_halo_kernel(const E4SConv p, const uint8_t* __restrict__ wpk, const int tiles_x, const int tiles_y, const int n_tiles,
                    const int total_jobs_in, const int4* __restrict__ rjobs, const int* __restrict__ rjob_count, const int bn_packed) {
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;   // device-side launch predicate (e4s_b200.h)
  constexpr int B_BYTES = BN * 128;               // one bf16 weight tile (hi or lo) of one phase of a packed 64-wide K chunk
  constexpr int P = UP ? 4 : 1;
  constexpr int PM = hl_phase_merge(BN, UP);      // phases merged into one MMA (N = PM * BN)
  constexpr int PL = P / PM;                      // MMA groups of PM merged phases
  constexpr int BST = hl_b_stages(BN, PM);
  constexpr int STAGE_B = 2 * PM * B_BYTES;       // [hi: PM x BN rows][lo: PM x BN rows]
  const uint32_t IDESC = umma_idesc_fmt(umma_idesc(PM * BN), p.tc_fmt);
  const bool f16 = p.tc_fmt == E4S_TC_F16;
  // same-resolution layers with BN <= 64: the hi and lo weight tiles of a stage are adjacent rows of one operand, so
  // A_hi x [B_hi ; B_lo] is ONE MMA with N = 2*BN (two accumulator halves, summed in the epilogue) followed by A_lo x B_hi:
  // 2 instead of 3 MMAs per K step (an M=128 MMA costs the same ~61 cycles for every N <= 128, profiles/r1_umma_rate_experiment.txt)
  constexpr bool NC = !UP && BN <= 64;
  const uint32_t IDESC2 = umma_idesc_fmt(umma_idesc(2 * BN <= 256 ? 2 * BN : 256), p.tc_fmt);
  // fp16 operands (same-resolution layers only): hi*hi accumulates alone, the two small terms go to a second accumulator (tc_ptx.cuh,
  // TC_LO_SCALE) -- for NC that is the second half the N-merged MMA already writes, otherwise BN more columns per set
  constexpr int ACC_COLS = P * BN * (UP ? 1 : 2);            // TMEM columns of one accumulator set
  constexpr int NSETS = (2 * ACC_COLS <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = NSETS * ACC_COLS <= 32 ? 32 : NSETS * ACC_COLS <= 64 ? 64 : NSETS * ACC_COLS <= 128 ? 128 : NSETS * ACC_COLS <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // [halo stage 0: hi | lo][halo stage 1: hi | lo][B ring: BST x (hi | lo)][barriers]
  constexpr int HALO_BYTES = 2 * HL_PLANE;
  constexpr int B_OFF = HL_HALO_STAGES * HALO_BYTES;
  float* s_epi = reinterpret_cast<float*>(smem + B_OFF + BST * STAGE_B);            // [2 slots][mul | add | slope | rgb0 | rgb1 | rgb2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF + BST * STAGE_B + 2 * 6 * BN * 4);
  const uint32_t bar_hfull = smem_u32(bars);                 // 2
  const uint32_t bar_hempty = bar_hfull + 16;                // 2
  const uint32_t bar_bfull = bar_hempty + 16;                // up to HL_MAX_BST
  const uint32_t bar_bempty = bar_bfull + 8 * HL_MAX_BST;
  const uint32_t bar_afull = bar_bempty + 8 * HL_MAX_BST;    // 2
  const uint32_t bar_aempty = bar_afull + 16;                // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * HL_MAX_BST);
  long long* s_acct = reinterpret_cast<long long*>(bars + 24);                     // [4 roles][8 laps]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dbg = g_hl_dbg;                                  // read ONCE (a global load per use otherwise)
  const bool acct = g_hl_trace != nullptr && blockIdx.x == 0 && (tid == 0 || tid == 8 * 32 || tid == HL_MMA_WARP * 32 || tid == (HL_MMA_WARP + 1) * 32);
  long long acct_t = 0;
  constexpr int cin_eff = C32 ? 32 : 64;                     // channels per halo row actually used (cin == 32 or cin % 64 == 0)
  const int G = C32 ? 1 : p.cin / 64;                        // 64-channel groups
  constexpr int ksteps = cin_eff / 16;
  constexpr int tpc = 64 / cin_eff;                          // taps per packed 64-wide K chunk (1, or 2 when cin == 32)
  constexpr int CPG = (9 + tpc - 1) / tpc;                   // packed chunks per (phase, group)
  const int num_kc = (9 * p.cin + 63) / 64;
  // masked layers: one job per (tile, region present in the tile) from a device-built list (e4s_region_tile_jobs);
  // the halo is modulated with that region's style and the epilogue keeps only the rows that belong to the region
  const int total_jobs = rjobs ? __ldg(rjob_count) * n_tiles : total_jobs_in;
  const int my_jobs = total_jobs > (int)blockIdx.x ? (total_jobs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // every weight chunk of a job fits the ring and all jobs use the same tiles: load once, keep resident
  const bool resident = n_tiles == 1 && G * PL * CPG <= BST;
  // Cluster launch (n_tiles == 1: every CTA walks the same weight chunk sequence): each CTA fetches 1/csz of every
  // stage from L2 and multicasts it to all CTAs of the cluster -> L2->SM weight traffic / csz (the 64..128-channel
  // layers re-stream 150 KB..1.1 MB of weights per 128-pixel job and were bound by exactly that traffic).  A stage is
  // refilled when EVERY CTA's MMAs have read it: tcgen05.commit multicast arrives on bempty of all CTAs (count = csz).
  const uint32_t csz = cluster_nctas(), crank = cluster_rank();
  const bool mcast = csz > 1 && !resident;
  const uint16_t cmask = (uint16_t)((1u << csz) - 1u);
  // all CTAs of a cluster must run the same number of weight-ring iterations: pad with weight-only jobs
  const int ring_jobs = mcast ? (total_jobs + (int)gridDim.x - 1) / (int)gridDim.x : my_jobs;
  const bool pow2 = ((tiles_x & (tiles_x - 1)) | (tiles_y & (tiles_y - 1))) == 0;
  const int sx_sh = 31 - __clz(tiles_x), sy_sh = 31 - __clz(tiles_y);

  auto decode = [&](int it) {
    int j = (int)blockIdx.x + it * (int)gridDim.x;
    HlJob r;
    if (n_tiles == 1) {
      r.nt = 0;
    } else {
      r.nt = j % n_tiles;
      j /= n_tiles;
    }
    if (rjobs) {
      const int4 jv = __ldg(rjobs + j);
      r.b = jv.x;
      r.y0 = jv.y;
      r.x0 = jv.z;
      r.reg = jv.w;
      return r;
    }
    r.reg = 0;
    if (pow2) {                                              // power-of-two grids: shifts instead of divisions
      r.x0 = (j & (tiles_x - 1)) * HL_TW;
      j >>= sx_sh;
      r.y0 = (j & (tiles_y - 1)) * HL_TH;
      r.b = j >> sy_sh;
      return r;
    }
    r.x0 = (j % tiles_x) * HL_TW;
    j /= tiles_x;
    r.y0 = (j % tiles_y) * HL_TH;
    r.b = j / tiles_y;
    return r;
  };

  if (tid < 32) s_acct[tid] = 0;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_hfull + 8 * s, TC_PRODUCER_WARPS);
      mbar_init(bar_hempty + 8 * s, 1);
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 4);
    }
    for (int s = 0; s < HL_MAX_BST; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, mcast ? csz : 1);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == HL_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (csz > 1) cluster_sync_all();                          // barrier inits visible cluster-wide before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (acct) acct_t = clock64();

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== halo producers ===================================================
    const int cgs = cin_eff == 32 ? 2 : 3;                   // lanes per pixel = cin_eff / 8 (4 or 8)
    const int cg = tid & ((1 << cgs) - 1);
    const int px0 = tid >> cgs;                              // halo pixels px0, px0 + pxs, ...
    const int pxs = 256 >> cgs;
    const int total_hg = my_jobs * G;                        // halo fills this CTA performs

    float4 v[HL_ITEMS][2];
    uint32_t okm = 0;
    float4 sc[2], mn[2], rs[2];                              // per-(sample, channel) modulation / InstanceNorm of this fill
    int pit = 0, pg = 0;                                     // (job, channel group) of the next prefetch
    auto prefetch = [&]() {
      okm = 0;
      if (pit >= my_jobs) return;
      const HlJob jb = decode(pit);
      const int ch = pg * 64 + cg * 8;
      if (++pg == G) {
        pg = 0;
        ++pit;
      }
      if (p.smod) {
        const float4* sp = reinterpret_cast<const float4*>(p.smod + ((int64_t)jb.b * p.regions + jb.reg) * p.cin + ch);
        sc[0] = __ldg(sp);
        sc[1] = __ldg(sp + 1);
      }
      if (p.in_mean) {
        const float4* mp = reinterpret_cast<const float4*>(p.in_mean + (int64_t)jb.b * p.cin + ch);
        const float4* qp = reinterpret_cast<const float4*>(p.in_rstd + (int64_t)jb.b * p.cin + ch);
        mn[0] = __ldg(mp); mn[1] = __ldg(mp + 1);
        rs[0] = __ldg(qp); rs[1] = __ldg(qp + 1);
      }
      const float* xb = p.x + (int64_t)jb.b * p.hin * p.win * p.x_pitch + ch;
#pragma unroll
      for (int i = 0; i < HL_ITEMS; ++i) {
        const int px = px0 + pxs * i;
        const int hy = px / HL_HP, hx = px - hy * HL_HP;
        const int iy = jb.y0 - 1 + hy, ix = jb.x0 - 1 + hx;
        if (px < HL_HPIX && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win) {
          okm |= 1u << i;
          const float4* src = reinterpret_cast<const float4*>(xb + (int64_t)(iy * p.win + ix) * p.x_pitch);
          v[i][0] = ldg_stream4(src);
          v[i][1] = ldg_stream4(src + 1);
        }
      }
    };

    if (!(dbg & 32)) prefetch();
    for (int hg = 0; hg < total_hg; ++hg) {
      const int hs = hg & 1;
      HL_LAP(0, 0);
      mbar_wait(bar_hempty + 8 * hs, ((hg >> 1) & 1) ^ 1);
      HL_LAP(0, 1);
      uint8_t* h_hi = smem + hs * HALO_BYTES;
      uint8_t* h_lo = h_hi + HL_PLANE;
      if (!(dbg & 32)) {
#pragma unroll
        for (int i = 0; i < HL_ITEMS; ++i) {
          const int px = px0 + pxs * i;
          if (px >= HL_HPIX) continue;
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (okm & (1u << i)) {
            f[0] = v[i][0].x; f[1] = v[i][0].y; f[2] = v[i][0].z; f[3] = v[i][0].w;
            f[4] = v[i][1].x; f[5] = v[i][1].y; f[6] = v[i][1].z; f[7] = v[i][1].w;
            if (p.in_mean) {
              f[0] = (f[0] - mn[0].x) * rs[0].x; f[1] = (f[1] - mn[0].y) * rs[0].y; f[2] = (f[2] - mn[0].z) * rs[0].z; f[3] = (f[3] - mn[0].w) * rs[0].w;
              f[4] = (f[4] - mn[1].x) * rs[1].x; f[5] = (f[5] - mn[1].y) * rs[1].y; f[6] = (f[6] - mn[1].z) * rs[1].z; f[7] = (f[7] - mn[1].w) * rs[1].w;
            }
            if (p.smod) {
              f[0] *= sc[0].x; f[1] *= sc[0].y; f[2] *= sc[0].z; f[3] *= sc[0].w;
              f[4] *= sc[1].x; f[5] *= sc[1].y; f[6] *= sc[1].z; f[7] *= sc[1].w;
            }
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) tc_split2(f16, f[2 * j], f[2 * j + 1], hi[j], lo[j]);
          // 128B swizzle on ABSOLUTE shared-memory address bits [7:9] (the plane base is only 128-byte aligned)
          const uint32_t rowaddr = (uint32_t)(hs * HALO_BYTES) + px * 128;
          const uint32_t off = px * 128 + ((cg ^ ((rowaddr >> 7) & 7)) << 4);
          const uint32_t off_lo = px * 128 + ((cg ^ (((rowaddr + HL_PLANE) >> 7) & 7)) << 4);
          *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(h_lo + off_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      HL_LAP(0, 2);
      // publish the halo FIRST: fence.proxy.async waits for this thread's outstanding memory operations, so issuing the
      // next job's global loads before it would serialise their full DRAM latency into every job
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfull + 8 * hs);
      HL_LAP(0, 3);
      if (!(dbg & 32)) prefetch();
      HL_LAP(0, 4);
    }
  } else if (warp < HL_MMA_WARP) {
    // =========================== epilogue warpgroup ===============================================
    // Software pipeline: the job descriptor is decoded two jobs ahead and the per-pixel operands (noise, float mask,
    // region label) are LOADED one job ahead and left untouched in registers until their job (any arithmetic on them
    // right after the load would stall this warp for the full L2 latency).  The fused per-channel vectors in shared
    // memory are rebuilt only when (sample, region, n-tile) changes.
    const int q = warp & 3;
    const int row = q * 32 + lane;                           // GEMM row = ty*8 + tx
    const int ty = row >> 3, tx = row & 7;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const bool need_lab = rjobs != nullptr;
    const bool noise1 = p.noise && p.noise_sc == 0;
    const bool fast = tc_epi_is_fast(p) && !(dbg & (1 | 2 | 8 | 16 | 128));    // bit7: force the generic epilogue
    const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
    // accumulate steps into the main accumulator: 9 taps x ksteps per group, 3 MMAs each (2 where hi|lo weights merge along N)
    const float corr = tc_acc_unbias(p, G * 9 * ksteps * (f16 ? 1 : (NC ? 2 : 3)));
    const float low = f16 ? 1.f / TC_LO_SCALE : 1.f;         // weight of the second accumulator half
    // fused ToRGB tail (host guarantees: !UP, one n-tile, no region jobs, fast epilogue)
    const bool do_rgb = !UP && p.rgb != nullptr;
    const int sk_h = p.hout >> 1, sk_w = p.wout >> 1;
    // FIR up-sampling of the skip image (upfirdn2d up=2, pad=(2,1), flipped 4x4 kernel): output (y, x) reads the 2x2 skip
    // pixels (sy0 + a, sx0 + d) with taps (ky0 + 2a, kx0 + 2d); tiles start at even coordinates, so the taps are per-thread constants
    const int ky0 = ty & 1, kx0 = tx & 1;
    float fk[4] = {0.f, 0.f, 0.f, 0.f};
    if (do_rgb && p.rgb_skip) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int d = 0; d < 2; ++d) fk[2 * a + d] = __ldg(p.rgb_fir + (3 - (ky0 + 2 * a)) * 4 + (3 - (kx0 + 2 * d)));
    }
    // per-thread constants of the skip taps: y0 / x0 are multiples of 16 / 8, so (oy0 - 2 + ky0) >> 1 = y0/2 + sk_dy
    const int sk_dy = (ty - 2 + ky0) >> 1, sk_dx = (tx - 2 + kx0) >> 1;
    const int sk_plane = sk_h * sk_w;
    const float rb0 = (do_rgb && p.rgb_bias) ? __ldg(p.rgb_bias) : 0.f, rb1 = (do_rgb && p.rgb_bias) ? __ldg(p.rgb_bias + 1) : 0.f,
                rb2 = (do_rgb && p.rgb_bias) ? __ldg(p.rgb_bias + 2) : 0.f;

    struct RowOps {
      float nz[P], pw[P];
      uint32_t lab[P];
    };
    auto fetch_rows = [&](const HlJob& j, RowOps& o) {
#pragma unroll
      for (int ph = 0; ph < P; ++ph) {
        o.nz[ph] = 0.f;
        o.pw[ph] = 1.f;
        o.lab[ph] = 0;
        const int oy = UP ? 2 * (j.y0 + ty) + (ph >> 1) : j.y0 + ty;
        const int ox = UP ? 2 * (j.x0 + tx) + (ph & 1) : j.x0 + tx;
        if (need_lab || p.pixw) {
          const int sy = nearest_src(oy, p.lab_h, p.hout), sx = nearest_src(ox, p.lab_w, p.wout);
          if (need_lab) o.lab[ph] = p.labels[((int64_t)j.b * p.lab_h + sy) * p.lab_w + sx];
          if (p.pixw) o.pw[ph] = __ldg(p.pixw + (int64_t)j.b * p.pixw_sb + (int64_t)sy * p.lab_w + sx);
        }
        if (noise1) o.nz[ph] = __ldg(p.noise + (int64_t)j.b * p.noise_sb + (int64_t)oy * p.wout + ox);
      }
    };

    HlJob jb = decode(0), jb1 = decode(my_jobs > 1 ? 1 : 0);
    RowOps cur, nxt;
    if (my_jobs > 0) fetch_rows(jb, cur);
    nxt = cur;
    int kb = -1, knt = -1, kreg = -1, slot = 0;
    for (int it = 0; it < my_jobs; ++it) {
      const int set = NSETS == 2 ? (it & 1) : 0;
      const int use = NSETS == 2 ? (it >> 1) : it;
      const HlJob jb2 = decode(it + 2 < my_jobs ? it + 2 : it);
      if (it + 1 < my_jobs) fetch_rows(jb1, nxt);
      const float* drow = p.demod ? p.demod + ((int64_t)jb.b * p.regions + jb.reg) * p.cout : nullptr;
      const int oy0 = UP ? 2 * (jb.y0 + ty) : jb.y0 + ty, ox0 = UP ? 2 * (jb.x0 + tx) : jb.x0 + tx;
      const int64_t pix0 = ((int64_t)jb.b * p.hout + oy0) * p.wout + ox0;
      HL_LAP(1, 0);
      if (jb.b != kb || jb.nt != knt || jb.reg != kreg) {
        // fused per-channel vectors of this (sample, region, n-tile) -> the other shared-memory slot.  Every warp has left
        // the jobs that read that slot before it passed the previous rebuild's barrier.
        kb = jb.b; knt = jb.nt; kreg = jb.reg;
        slot ^= 1;
        float* svw = s_epi + slot * 6 * BN;
        for (int n = tid - 8 * 32; n < BN; n += 128) {
          const int ng = jb.nt * BN + n;
          float mul = drow ? __ldg(drow + ng) : 1.f;
          if (p.ch_scale) mul *= __ldg(p.ch_scale + ng);
          svw[n] = mul * (fast ? corr : 1.f);            // the generic path scales the accumulator itself
          svw[BN + n] = p.ch_shift ? __ldg(p.ch_shift + ng) : 0.f;
          svw[2 * BN + n] = tc_epi_slope(p, ng);
          if (do_rgb) {
            const float sm = __ldg(p.rgb_smod + (int64_t)jb.b * p.cout + ng);
#pragma unroll
            for (int c = 0; c < 3; ++c) svw[(3 + c) * BN + n] = ng < p.cout ? __ldg(p.rgb_w + c * p.cout + ng) * sm : 0.f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps only
      }
      const float* sv = s_epi + slot * 6 * BN;
      // skip taps of this pixel: issued before the accumulator wait, consumed (first arithmetic) after the channel loop
      float skv[12];
      if (do_rgb && p.rgb_skip) {
        const int sy0 = (jb.y0 >> 1) + sk_dy, sx0 = (jb.x0 >> 1) + sk_dx;     // -1 at the top / left border
        const bool vy[2] = {sy0 >= 0, sy0 + 1 < sk_h}, vx[2] = {sx0 >= 0, sx0 + 1 < sk_w};
        const float* sp = p.rgb_skip + (int64_t)jb.b * 3 * sk_plane + (sy0 * sk_w + sx0);   // 32-bit offsets from here on
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int d = 0; d < 2; ++d) skv[c * 4 + 2 * a + d] = (vy[a] && vx[d]) ? __ldg(sp + (c * sk_plane + a * sk_w + d)) : 0.f;
      }
      HL_LAP(1, 3);
      mbar_wait(bar_afull + 8 * set, use & 1);
      tc_fence_after();
      HL_LAP(1, 1);
#pragma unroll
      for (int ph = 0; ph < P; ++ph) {
        const uint32_t tacc = tmem_base + (uint32_t)(set * ACC_COLS + ph * BN) + ((uint32_t)(q * 32) << 16);
        const int64_t pix = UP ? pix0 + (ph >> 1) * p.wout + (ph & 1) : pix0;
        const bool live = !need_lab || cur.lab[ph] == (uint32_t)jb.reg;
        if (fast) {
          float* optr = p.out ? p.out + pix * p.out_pitch + jb.nt * BN : nullptr;
          const float nz = nw * cur.nz[ph];
          float rgb3[3] = {0.f, 0.f, 0.f};
          if (NC || (!UP && f16)) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
              uint32_t ra[16], rb[16];
              float acc[16];
              tmem_ld16_nowait(tacc + (uint32_t)c0, ra);
              tmem_ld16_nowait(tacc + (uint32_t)(BN + c0), rb);
              tmem_wait_ld16x2(ra, rb);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[j] = fmaf(__uint_as_float(rb[j]), low, __uint_as_float(ra[j]));
              if (live) {
                if (do_rgb) tc_epilogue_fast<16, true>(optr ? optr + c0 : nullptr, acc, sv, c0, BN, nz, gain, sv + 3 * BN, rgb3);
                else tc_epilogue_fast<16>(optr + c0, acc, sv, c0, BN, nz, gain);
              }
            }
          } else {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
              uint32_t ra[32];
              float acc[32];
              tmem_ld32_nowait(tacc + (uint32_t)c0, ra);
              tmem_wait_ld32(ra);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(ra[j]);
              if (live) {
                if (do_rgb) tc_epilogue_fast<32, true>(optr ? optr + c0 : nullptr, acc, sv, c0, BN, nz, gain, sv + 3 * BN, rgb3);
                else tc_epilogue_fast<32>(optr + c0, acc, sv, c0, BN, nz, gain);
              }
            }
          }
          if (do_rgb) {
            const int64_t hw = (int64_t)p.hout * p.wout;
            float* ro = p.rgb + (int64_t)jb.b * 3 * hw + (int64_t)oy0 * p.wout + ox0;
            const float bias3[3] = {rb0, rb1, rb2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float v = rgb3[c] + bias3[c];
              if (p.rgb_skip) v += fk[0] * skv[c * 4] + fk[1] * skv[c * 4 + 1] + fk[2] * skv[c * 4 + 2] + fk[3] * skv[c * 4 + 3];
              ro[c * hw] = v;
            }
          }
          continue;
        }
        TcEpiRow er;
        er.pix = pix;
        er.drow = drow;
        er.pw = cur.pw[ph];
        er.nw = nw;
        er.nrow = p.noise ? p.noise + (int64_t)jb.b * p.noise_sb + (int64_t)(UP ? oy0 + (ph >> 1) : oy0) * p.wout + (UP ? ox0 + (ph & 1) : ox0) : nullptr;
        er.nz = nw * cur.nz[ph];
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float acc[16];
          if (!(dbg & 2)) {
            tmem_ld16(tacc + (uint32_t)c0, acc);
            if (NC || (!UP && f16)) {
              float acc2[16];
              tmem_ld16(tacc + (uint32_t)(BN + c0), acc2);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc2[j], low, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] *= corr;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
          }
          if (!(dbg & 1) && live) tc_epilogue16_sv(p, acc, jb.nt * BN + c0, c0, BN, sv, er, dbg);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aempty + 8 * set);
      HL_LAP(1, 2);
      jb = jb1;
      jb1 = jb2;
      cur = nxt;
    }
  } else if (warp == HL_MMA_WARP) {
    // =========================== MMA issuer =========================================================
    // ONE elected thread runs the whole role: the instruction stream of this warp is the critical path of the kernel
    // (profiles/r1_halo_accounting_*.txt: ~450 cycles of loop overhead per weight chunk before this rewrite), so the tap /
    // K-step loops are unrolled with compile-time descriptor offsets and there is no per-chunk elect / syncwarp.
    if (elect_one()) {
      constexpr uint32_t A_HI = umma_desc_hi(HL_HP * 128), B_HI = umma_desc_hi(1024);
      constexpr uint32_t kb16 = (uint32_t)(cin_eff * 2) >> 4;   // 16-byte units per tap inside a packed 64-wide K chunk
      const uint32_t b_ring = umma_desc_lo(smem_base + B_OFF);
      const uint32_t lo_col = f16 ? (uint32_t)BN : 0u;
      int hg = 0, bs = 0, bph = 0;                             // running halo-fill counter, weight stage and its phase parity
      for (int it = 0; it < my_jobs; ++it) {
        if (resident) bs = 0;                                  // resident weights: chunk i of every job lives in stage i
        const int set = NSETS == 2 ? (it & 1) : 0;
        const int use = NSETS == 2 ? (it >> 1) : it;
        HL_LAP(2, 0);
        mbar_wait(bar_aempty + 8 * set, (use & 1) ^ 1);
        tc_fence_after();
        HL_LAP(2, 1);
        for (int g = 0; g < G; ++g, ++hg) {
          const int hs = hg & 1;
          mbar_wait(bar_hfull + 8 * hs, (hg >> 1) & 1);
          HL_LAP(2, 2);
          tc_fence_after();
          const uint32_t a_h = umma_desc_lo(smem_base + hs * HALO_BYTES), a_l = a_h + (HL_PLANE >> 4);
#pragma unroll
          for (int pl = 0; pl < PL; ++pl) {
            const uint32_t tacc = tmem_base + (uint32_t)(set * ACC_COLS + pl * PM * BN);
#pragma unroll
            for (int c = 0; c < CPG; ++c) {
              HL_LAP(2, 5);
              if (!resident || it == 0) {
                mbar_wait(bar_bfull + 8 * bs, bph);
                tc_fence_after();
              }
              HL_LAP(2, 4);
              const uint32_t b_h = b_ring + (uint32_t)bs * (STAGE_B >> 4), b_l = b_h + ((PM * B_BYTES) >> 4);
              if (!(dbg & 4)) {
#pragma unroll
                for (int tt = 0; tt < tpc; ++tt) {
                  const int t9 = c * tpc + tt;
                  if (t9 < 9) {
                    const uint32_t aoff = (uint32_t)((t9 / 3) * HL_HP + (t9 % 3)) * 8u;
                    const uint32_t boff = (uint32_t)tt * kb16;
#pragma unroll
                    for (int k = 0; k < ksteps; ++k) {
                      const uint32_t ao = aoff + 2 * k, bo = boff + 2 * k;
                      const uint32_t first = (t9 | k) != 0 ? 1u : (uint32_t)(g != 0);
                      if (NC) {
                        umma_bf16_w(tacc, a_h + ao, A_HI, b_h + bo, B_HI, IDESC2, first);   // A_hi x [B_hi ; B_lo] -> both accumulator halves
                        umma_bf16_w(tacc + lo_col, a_l + ao, A_HI, b_h + bo, B_HI, IDESC, 1);   // A_lo x B_hi -> first half (fp16: second)
                      } else if (!UP && f16) {                                              // separate small-term accumulator
                        umma_bf16_w(tacc + BN, a_l + ao, A_HI, b_h + bo, B_HI, IDESC, first);
                        umma_bf16_w(tacc + BN, a_h + ao, A_HI, b_l + bo, B_HI, IDESC, 1);
                        umma_bf16_w(tacc, a_h + ao, A_HI, b_h + bo, B_HI, IDESC, first);
                      } else {
                        umma_bf16_w(tacc, a_l + ao, A_HI, b_h + bo, B_HI, IDESC, first);
                        umma_bf16_w(tacc, a_h + ao, A_HI, b_l + bo, B_HI, IDESC, 1);
                        umma_bf16_w(tacc, a_h + ao, A_HI, b_h + bo, B_HI, IDESC, 1);
                      }
                    }
                  }
                }
              }
              if (!resident) {
                if (mcast) umma_commit_mc(bar_bempty + 8 * bs, cmask);
                else umma_commit(bar_bempty + 8 * bs);
              }
              if (++bs == BST) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          umma_commit(bar_hempty + 8 * hs);                    // every tap of every phase has read this halo
        }
        umma_commit(bar_afull + 8 * set);
        HL_LAP(2, 3);
      }
      // weight-only padding jobs: keep this CTA's share of the cluster's weight ring turning (release every stage unread)
      for (int it = my_jobs; it < ring_jobs; ++it)
        for (int c = 0; c < G * PL * CPG; ++c) {
          mbar_wait(bar_bfull + 8 * bs, bph);
          tc_fence_after();
          umma_commit_mc(bar_bempty + 8 * bs, cmask);
          if (++bs == BST) {
            bs = 0;
            bph ^= 1;
          }
        }
    }
    __syncwarp();
  } else {
    // =========================== weight loader (one elected thread) ================================
    if (elect_one()) {
      int bs = 0, bph = 0;
      for (int it = 0; it < (resident ? (my_jobs > 0 ? 1 : 0) : ring_jobs); ++it) {
        HlJob jb = decode(it < my_jobs ? it : 0);
        if (mcast) jb.nt = 0;
        for (int g = 0; g < G; ++g)
          for (int pl = 0; pl < PL; ++pl)
            for (int c = 0; c < CPG; ++c) {
              const int kc = tpc == 1 ? g * 9 + c : c;       // packed chunk order: channel group outer, tap inner (pack_weights_tc)
              HL_LAP(3, 1);
              mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
              HL_LAP(3, 0);
              if (dbg & 64) {
                mbar_arrive(bar_bfull + 8 * bs);
              } else {
                const uint32_t dst = smem_base + B_OFF + bs * STAGE_B;
                mbar_arrive_expect_tx(bar_bfull + 8 * bs, STAGE_B);
                // packed chunk (e4s_pack_weights_tc): [hi tiles of the P phases | lo tiles of the P phases], bn_packed rows each;
                // BN < bn_packed: this job's n-tile is a 128-row slice of a packed 256-row tile
                const int64_t tb = (int64_t)bn_packed * 128;
                const int ch0 = jb.nt * BN;
                const uint8_t* src = wpk + ((int64_t)(ch0 / bn_packed) * num_kc + kc) * (2 * (int64_t)P * tb) + (int64_t)(ch0 % bn_packed) * 128;
                if (bn_packed != BN) {                       // sliced tiles: hi and lo of every merged phase separately
#pragma unroll
                  for (int q = 0; q < PM; ++q) {
                    bulk_g2s(dst + q * B_BYTES, src + (int64_t)(pl * PM + q) * tb, B_BYTES, bar_bfull + 8 * bs);
                    bulk_g2s(dst + (PM + q) * B_BYTES, src + (int64_t)(P + pl * PM + q) * tb, B_BYTES, bar_bfull + 8 * bs);
                  }
                } else if (mcast) {                          // this CTA's slice of the stage, to every CTA of the cluster
                  const uint32_t slice = (uint32_t)STAGE_B / csz;
                  if (PM == P) {
                    bulk_g2s_mc(dst + crank * slice, src + crank * slice, slice, bar_bfull + 8 * bs, cmask);
                  } else {                                   // stage = [hi tile | lo tile] of phase pl: a slice never straddles the two
                    const uint32_t off = crank * slice;
                    const uint8_t* sp = off < (uint32_t)B_BYTES ? src + (int64_t)pl * B_BYTES + off : src + (int64_t)(P + pl) * B_BYTES + (off - B_BYTES);
                    bulk_g2s_mc(dst + off, sp, slice, bar_bfull + 8 * bs, cmask);
                  }
                } else if (PM == P) {                        // every phase merged (or a plain conv): the stage IS the chunk
                  bulk_g2s(dst, src, STAGE_B, bar_bfull + 8 * bs);
                } else {                                     // one phase per MMA group
                  bulk_g2s(dst, src + (int64_t)pl * B_BYTES, B_BYTES, bar_bfull + 8 * bs);
                  bulk_g2s(dst + B_BYTES, src + (int64_t)(P + pl) * B_BYTES, B_BYTES, bar_bfull + 8 * bs);
                }
              }
              if (++bs == BST) {
                bs = 0;
                bph ^= 1;
              }
            }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (csz > 1) cluster_sync_all();                          // no CTA may exit while a peer can still multicast into it
  if (warp == HL_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (g_hl_trace != nullptr && blockIdx.x == 0 && tid < 32 && tid < g_hl_trace_cap) g_hl_trace[tid] = (unsigned long long)s_acct[tid];
  if (g_hl_trace != nullptr && blockIdx.x == 0 && tid == 32 && 32 < g_hl_trace_cap) g_hl_trace[32] = (unsigned long long)my_jobs;
}

// One CTA per 16x8 tile: which regions own at least one of the tile's output pixels?  Appends one int4 job
// (b, y0, x0, region) per region present.  up2: the tile is over INPUT pixels and covers 32x16 output pixels.
__global__ void __launch_bounds__(128) region_tile_jobs_kernel(const uint8_t* __restrict__ labels, int lab_h, int lab_w, int hout,
                                                               int wout, int up2, int tiles_x, int tiles_y, int4* __restrict__ jobs,
                                                               int* __restrict__ count, int max_jobs) {
  __shared__ uint32_t present;
  if (threadIdx.x == 0) present = 0;
  __syncthreads();
  int t = blockIdx.x;
  const int txi = t % tiles_x;
  t /= tiles_x;
  const int tyi = t % tiles_y;
  const int b = t / tiles_y;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  uint32_t mine = 0;
  const int np = up2 ? 4 : 1;
  for (int ph = 0; ph < np; ++ph) {
    const int oy = up2 ? 2 * (tyi * HL_TH + ty) + (ph >> 1) : tyi * HL_TH + ty;
    const int ox = up2 ? 2 * (txi * HL_TW + tx) + (ph & 1) : txi * HL_TW + tx;
    const int sy = nearest_src(oy, lab_h, hout), sx = nearest_src(ox, lab_w, wout);
    mine |= 1u << (labels[((int64_t)b * lab_h + sy) * lab_w + sx] & 31);
  }
  atomicOr(&present, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = present;
    const int n = __popc(m);
    const int base = atomicAdd(count, n);
    int i = 0;
    while (m) {
      const int r = __ffs(m) - 1;
      m &= m - 1;
      if (base + i < max_jobs) jobs[base + i] = make_int4(b, tyi * HL_TH, txi * HL_TW, r);
      ++i;
    }
  }
}

static int g_halo_sm_count_dev[E4S_MAX_DEVICES] = {};
static int g_halo_cluster = -1;

// geometry the halo kernel takes (everything else stays on the gather kernel)
bool tc_halo_geometry_ok(const E4SConv* p) {
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  if (!up && !(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && p->in_shift == 0)) return false;
  if (p->hin % HL_TH || p->win % HL_TW) return false;
  if (!((p->cin == 32 && p->cout == 32) || p->cin % 64 == 0)) return false;   // cin == 32 is instantiated for the 32 -> 32 layers only
  const int bn = tc_block_n(p->cout);
  if ((up ? 4 : 1) * bn > 512) return false;
  return true;
}

bool tc_halo_eligible(const E4SConv* p) {
  if (p->labels) return false;                               // per-pixel regions: region-job list (e4s_conv_tc_regions) or gather
  return tc_halo_geometry_ok(p);
}

template <int BN, bool UP, bool C32 = false>
static int launch_halo(const E4SConv* p, const void* wpk, cudaStream_t s, const int4* rjobs = nullptr, const int* rjob_count = nullptr,
                       int rjob_host_count = 0) {
  static bool attr_set_dev[E4S_MAX_DEVICES] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  constexpr int pm = hl_phase_merge(BN, UP);
  constexpr int smem_bytes = hl_smem_bytes(BN, pm);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel<BN, UP, C32>, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_MAX);
    if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "conv_tc(halo): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int& g_halo_sm_count = g_halo_sm_count_dev[current_device_slot()];
  if (g_halo_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_halo_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_halo_sm_count <= 0) g_halo_sm_count = 148;
  }
  const int tiles_x = p->win / HL_TW, tiles_y = p->hin / HL_TH, n_tiles = p->cout / BN;
  const int bn_packed = tc_block_n(p->cout);                 // row count of the packed weight tiles (e4s_pack_weights_tc)
  const int64_t total = rjobs ? (int64_t)rjob_host_count * n_tiles : (int64_t)p->batch * tiles_x * tiles_y * n_tiles;
  E4S_REQUIRE(total > 0 && total < 0x7fffffff, "conv_tc(halo): bad job count");
  const unsigned grid = (unsigned)(total < g_halo_sm_count ? total : g_halo_sm_count);
  // cluster size for the weight multicast (E4S_HALO_CLUSTER = 1 | 2 | 4): full persistent grids with one n-tile only.
  // Default 1: measured on B200 the multicast changes nothing (the layers are bound by shared-memory bandwidth inside the
  // SM, not by L2->SM traffic), so the plain launch is kept as the production path.
  if (g_halo_cluster < 0) {
    const char* e = getenv("E4S_HALO_CLUSTER");
    g_halo_cluster = e ? atoi(e) : 1;
    if (g_halo_cluster != 1 && g_halo_cluster != 2 && g_halo_cluster != 4) g_halo_cluster = 1;
  }
  unsigned csz = (n_tiles == 1 && (int)grid == g_halo_sm_count && grid % (unsigned)g_halo_cluster == 0) ? (unsigned)g_halo_cluster : 1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(HL_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (csz > 1) {                                             // all clusters must be co-resident (persistent kernel)
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, conv_tc_halo_kernel<BN, UP, C32>, &cfg);
    if (e != cudaSuccess || max_clusters * (int)csz < (int)grid) {
      (void)cudaGetLastError();
      csz = 1;
      attr[0].val.clusterDim.x = 1;
    }
  }
  const uint8_t* wp = static_cast<const uint8_t*>(wpk);
  const int total_i = (int)total;
  cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc_halo_kernel<BN, UP, C32>, *p, wp, tiles_x, tiles_y, n_tiles, total_i, rjobs, rjob_count, bn_packed);
  if (le != cudaSuccess) return fail(E4S_ERR_CUDA, "e4s_conv_tc(halo): launch: %s", cudaGetErrorString(le));
  return check_launch("e4s_conv_tc(halo)");
}

int tc_halo_set_flags(int flags) {
  cudaError_t e = cudaMemcpyToSymbol(g_hl_dbg, &flags, sizeof(int));
  return e == cudaSuccess ? E4S_OK : fail(E4S_ERR_CUDA, "halo flags: %s", cudaGetErrorString(e));
}

int tc_halo_set_trace(void* buf, int cap_records) {
  unsigned long long* ptr = static_cast<unsigned long long*>(buf);
  cudaError_t e = cudaMemcpyToSymbol(g_hl_trace, &ptr, sizeof(ptr));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_hl_trace_cap, &cap_records, sizeof(int));
  return e == cudaSuccess ? E4S_OK : fail(E4S_ERR_CUDA, "halo trace: %s", cudaGetErrorString(e));
}

int tc_launch_halo(const E4SConv* p, const void* wpk, cudaStream_t s, const int4* rjobs, const int* rjob_count, int rjob_host_count) {
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  switch (tc_block_n(p->cout)) {
    // cout >= 256: 128-column n-tiles (slices of the packed 256-row tiles).  An M=128 MMA is operand-fetch bound up to
    // N = 128 (64 cycles) and compute bound at N = 256 (128 cycles): same columns per cycle, but a 64 KB N=256 weight
    // stage leaves room for ONE stage next to the halo buffers (0.28 ms per 512->512 @32^2 encoder conv, 40 % of its floor)
    case 256: return launch_halo<128, false>(p, wpk, s, rjobs, rjob_count, rjob_host_count);     // up: 4 x 256 columns exceed TMEM (geometry_ok)
    case 128: return up ? launch_halo<128, true>(p, wpk, s, rjobs, rjob_count, rjob_host_count) : launch_halo<128, false>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    case 64: return up ? launch_halo<64, true>(p, wpk, s, rjobs, rjob_count, rjob_host_count) : launch_halo<64, false>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    case 32:
      if (p->cin == 32)
        return up ? launch_halo<32, true, true>(p, wpk, s, rjobs, rjob_count, rjob_host_count) : launch_halo<32, false, true>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
      return up ? launch_halo<32, true>(p, wpk, s, rjobs, rjob_count, rjob_host_count) : launch_halo<32, false>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    default: return fail(E4S_ERR_UNSUPPORTED, "conv_tc(halo): unsupported cout %d", p->cout);
  }
}

}  // namespace e4s

extern "C" int e4s_region_tile_jobs(const uint8_t* labels, int batch, int lab_h, int lab_w, int hout, int wout, int up2, int32_t* jobs,
                                    int32_t* count, int max_jobs, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(labels && jobs && count && batch > 0 && lab_h > 0 && lab_w > 0 && hout > 0 && wout > 0 && max_jobs > 0, "region_tile_jobs: bad args");
  const int gh = up2 ? hout / 2 : hout, gw = up2 ? wout / 2 : wout;
  E4S_REQUIRE(gh % HL_TH == 0 && gw % HL_TW == 0 && (!up2 || (hout % 2 == 0 && wout % 2 == 0)), "region_tile_jobs: %dx%d is not tileable by 16x8", gh, gw);
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(jobs) & 15) == 0, "region_tile_jobs: jobs must be 16-byte aligned");
  const int tiles_x = gw / HL_TW, tiles_y = gh / HL_TH;
  region_tile_jobs_kernel<<<(unsigned)(batch * tiles_x * tiles_y), 128, 0, as_stream(stream)>>>(labels, lab_h, lab_w, hout, wout, up2, tiles_x,
                                                                                              tiles_y, reinterpret_cast<int4*>(jobs), count, max_jobs);
  return check_launch("region_tile_jobs");
}

// debug aid (not part of the reference-facing surface): install / remove a clock64 timeline buffer for CTA 0 of the halo kernel
extern "C" int e4s_debug_halo_trace(void* buf, int cap_records) { return e4s::tc_halo_set_trace(buf, cap_records); }
extern "C" int e4s_debug_halo_flags(int flags) { return e4s::tc_halo_set_flags(flags); }

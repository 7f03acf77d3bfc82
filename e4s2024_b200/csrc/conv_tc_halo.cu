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
__host__ __device__ constexpr int hl_fixed_bytes(int bn) { return HL_HALO_STAGES * 2 * HL_PLANE + 2 * 3 * bn * 4 + 512 + 1024; }
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

// optional timeline trace of CTA 0 (debug / profiling aid): records (role, job, event, clock64) when a buffer is installed
__device__ unsigned long long* g_hl_trace = nullptr;
__device__ int g_hl_trace_cap = 0;
__device__ int g_hl_trace_n = 0;
__device__ int g_hl_dbg = 0;      // profiling experiments: bit0 skip epilogue math+stores, bit1 skip tcgen05.ld, bit2 skip MMAs
__device__ __forceinline__ void hl_trace(int role, int it, int ev) {
  if (g_hl_trace == nullptr || blockIdx.x != 0) return;
  const int i = atomicAdd(&g_hl_trace_n, 1);
  if (i < g_hl_trace_cap) {
    g_hl_trace[2 * i] = ((unsigned long long)role << 48) | ((unsigned long long)(it & 0xffffff) << 16) | (unsigned long long)ev;
    g_hl_trace[2 * i + 1] = clock64();
  }
}

template <int BN>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_tc_halo_kernel(const E4SConv p, const uint8_t* __restrict__ wpk, const int tiles_x, const int tiles_y, const int n_tiles,
                    const int total_jobs_in, const int4* __restrict__ rjobs, const int* __restrict__ rjob_count) {
  constexpr int B_BYTES = BN * 128;               // one bf16 weight tile (hi or lo) of one phase of a packed 64-wide K chunk
  const bool up = p.mode == E4S_CONV_UP2_POLYPHASE;
  const int PM = hl_phase_merge(BN, up);          // phases merged into one MMA (N = PM * BN)
  const int BST = hl_b_stages(BN, PM);
  const int STAGE_B = 2 * PM * B_BYTES;           // [hi: PM x BN rows][lo: PM x BN rows]
  const uint32_t IDESC = umma_idesc(PM * BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // [halo stage 0: hi | lo][halo stage 1: hi | lo][B ring: BST x (hi | lo)][barriers]
  constexpr int HALO_BYTES = 2 * HL_PLANE;
  constexpr int B_OFF = HL_HALO_STAGES * HALO_BYTES;
  float* s_epi = reinterpret_cast<float*>(smem + B_OFF + BST * STAGE_B);            // [2 slots][mul | add | prelu][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF + BST * STAGE_B + 2 * 3 * BN * 4);
  const uint32_t bar_hfull = smem_u32(bars);                 // 2
  const uint32_t bar_hempty = bar_hfull + 16;                // 2
  const uint32_t bar_bfull = bar_hempty + 16;                // up to HL_MAX_BST
  const uint32_t bar_bempty = bar_bfull + 8 * HL_MAX_BST;
  const uint32_t bar_afull = bar_bempty + 8 * HL_MAX_BST;    // 2
  const uint32_t bar_aempty = bar_afull + 16;                // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * HL_MAX_BST);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = up ? 4 : 1;
  const int PL = P / PM;                                      // MMA groups of PM merged phases
  const int cin_eff = p.cin < 64 ? p.cin : 64;               // channels per halo row actually used
  const int G = p.cin < 64 ? 1 : p.cin / 64;                 // 64-channel groups
  const int ksteps = cin_eff / 16;
  const int tpc = 64 / cin_eff;                              // taps per packed 64-wide K chunk (1, or 2 when cin == 32)
  const int CPG = (9 + tpc - 1) / tpc;                       // packed chunks per (phase, group)
  const int num_kc = (9 * p.cin + 63) / 64;
  const int nsets = (2 * P * BN <= 512) ? 2 : 1;
  // masked layers: one job per (tile, region present in the tile) from a device-built list (e4s_region_tile_jobs);
  // the halo is modulated with that region's style and the epilogue keeps only the rows that belong to the region
  const int total_jobs = rjobs ? __ldg(rjob_count) * n_tiles : total_jobs_in;
  const int my_jobs = total_jobs > (int)blockIdx.x ? (total_jobs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  auto decode = [&](int it) {
    int j = (int)blockIdx.x + it * (int)gridDim.x;
    HlJob r;
    r.nt = j % n_tiles;
    j /= n_tiles;
    if (rjobs) {
      const int4 jv = __ldg(rjobs + j);
      r.b = jv.x;
      r.y0 = jv.y;
      r.x0 = jv.z;
      r.reg = jv.w;
      return r;
    }
    r.reg = 0;
    if (((tiles_x & (tiles_x - 1)) | (tiles_y & (tiles_y - 1))) == 0) {   // power-of-two grids: shifts instead of divisions
      const int sx = 31 - __clz(tiles_x), sy = 31 - __clz(tiles_y);
      r.x0 = (j & (tiles_x - 1)) * HL_TW;
      j >>= sx;
      r.y0 = (j & (tiles_y - 1)) * HL_TH;
      r.b = j >> sy;
      return r;
    }
    r.x0 = (j % tiles_x) * HL_TW;
    j /= tiles_x;
    r.y0 = (j % tiles_y) * HL_TH;
    r.b = j / tiles_y;
    return r;
  };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_hfull + 8 * s, TC_PRODUCER_WARPS);
      mbar_init(bar_hempty + 8 * s, 1);
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 4);
    }
    for (int s = 0; s < HL_MAX_BST; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(nsets * P * BN)) tmem_cols <<= 1;
  if (warp == HL_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== halo producers ===================================================
    const int cg = tid & 7;
    const int px0 = tid >> 3;                                // halo pixels px0, px0+32, ...
    const bool cg_live = cg * 8 < cin_eff;
    const int total_hg = my_jobs * G;                        // halo fills this CTA performs

    float4 v[HL_ITEMS][2];
    uint32_t okm = 0;
    float4 sc[2], mn[2], rs[2];                              // per-(sample, channel) modulation / InstanceNorm of this fill
    auto prefetch = [&](int hg) {
      okm = 0;
      if (hg >= total_hg || !cg_live) return;
      const int it = hg / G, g = hg - it * G;
      const HlJob jb = decode(it);
      const int ch = g * 64 + cg * 8;
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
#pragma unroll
      for (int i = 0; i < HL_ITEMS; ++i) {
        const int px = px0 + 32 * i;
        const int hy = px / HL_HP, hx = px - hy * HL_HP;
        const int iy = jb.y0 - 1 + hy, ix = jb.x0 - 1 + hx;
        if (px < HL_HPIX && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win) {
          okm |= 1u << i;
          const float4* src = reinterpret_cast<const float4*>(p.x + (((int64_t)jb.b * p.hin + iy) * p.win + ix) * p.x_pitch + ch);
          v[i][0] = ldg_stream4(src);
          v[i][1] = ldg_stream4(src + 1);
        }
      }
    };

    const int pdbg = g_hl_dbg;
    if (!(pdbg & 32)) prefetch(0);
    for (int hg = 0; hg < total_hg; ++hg) {
      const int hs = hg & 1;
      if (tid == 0) hl_trace(0, hg, 0);
      mbar_wait(bar_hempty + 8 * hs, ((hg >> 1) & 1) ^ 1);
      if (tid == 0) hl_trace(0, hg, 1);
      uint8_t* h_hi = smem + hs * HALO_BYTES;
      uint8_t* h_lo = h_hi + HL_PLANE;
      if (cg_live && !(pdbg & 32)) {
#pragma unroll
        for (int i = 0; i < HL_ITEMS; ++i) {
          const int px = px0 + 32 * i;
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
          for (int j = 0; j < 4; ++j) {
            const float a = f[2 * j], b = f[2 * j + 1];
            const uint32_t h = pack_bf16x2(a, b);
            hi[j] = h;
            lo[j] = pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xffff0000u));
          }
          // 128B swizzle on ABSOLUTE shared-memory address bits [7:9] (the plane base is only 128-byte aligned)
          const uint32_t rowaddr = (uint32_t)(hs * HALO_BYTES) + px * 128;
          const uint32_t off = px * 128 + ((cg ^ ((rowaddr >> 7) & 7)) << 4);
          const uint32_t off_lo = px * 128 + ((cg ^ (((rowaddr + HL_PLANE) >> 7) & 7)) << 4);
          *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(h_lo + off_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      if (tid == 0) hl_trace(0, hg, 2);
      // publish the halo FIRST: fence.proxy.async waits for this thread's outstanding memory operations, so issuing the
      // next job's global loads before it would serialise their full DRAM latency into every job
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfull + 8 * hs);
      if (!(pdbg & 32)) prefetch(hg + 1);
      if (tid == 0) hl_trace(0, hg, 3);
    }
  } else if (warp < HL_MMA_WARP) {
    // =========================== epilogue warpgroup ===============================================
    const int q = warp & 3;
    const int row = q * 32 + lane;                           // GEMM row = ty*8 + tx
    const int ty = row >> 3, tx = row & 7;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const int dbg = g_hl_dbg;
    for (int it = 0; it < my_jobs; ++it) {
      const HlJob jb = decode(it);
      const int set = nsets == 2 ? (it & 1) : 0;
      const int use = nsets == 2 ? (it >> 1) : it;
      // per-pixel epilogue operands (noise / float-mask loads) are fetched BEFORE waiting for the accumulator
      const float* drow = p.demod ? p.demod + ((int64_t)jb.b * p.regions + jb.reg) * p.cout : nullptr;
      TcEpiRow er[4];
      uint32_t rowlive = 0xf;                                   // per phase: does this row (pixel) belong to the job's region
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        if (ph >= P) break;
        const int oy = up ? 2 * (jb.y0 + ty) + (ph >> 1) : jb.y0 + ty;
        const int ox = up ? 2 * (jb.x0 + tx) + (ph & 1) : jb.x0 + tx;
        er[ph].pix = ((int64_t)jb.b * p.hout + oy) * p.wout + ox;
        if (rjobs) {
          const int sy = nearest_src(oy, p.lab_h, p.hout), sx = nearest_src(ox, p.lab_w, p.wout);
          if (p.labels[((int64_t)jb.b * p.lab_h + sy) * p.lab_w + sx] != jb.reg) rowlive &= ~(1u << ph);
        }
        er[ph].drow = drow;
        er[ph].pw = 1.f;
        if (p.pixw) {
          const int sy = nearest_src(oy, p.lab_h, p.hout), sx = nearest_src(ox, p.lab_w, p.wout);
          er[ph].pw = __ldg(p.pixw + (int64_t)jb.b * p.pixw_sb + (int64_t)sy * p.lab_w + sx);
        }
        er[ph].nw = nw;
        er[ph].nrow = p.noise ? p.noise + (int64_t)jb.b * p.noise_sb + (int64_t)oy * p.wout + ox : nullptr;
        er[ph].nz = (er[ph].nrow && p.noise_sc == 0) ? nw * __ldg(er[ph].nrow) : 0.f;
      }
      if (warp == 8 && lane == 0) hl_trace(1, it, 0);
      // fused per-channel vectors of this job -> shared memory (demod is per sample; everything else per layer)
      float* sv = s_epi + (it & 1) * 3 * BN;
      for (int n = tid - 8 * 32; n < BN; n += 128) {
        const int ng = jb.nt * BN + n;
        float mul = drow ? __ldg(drow + ng) : 1.f;
        if (p.ch_scale) mul *= __ldg(p.ch_scale + ng);
        sv[n] = mul;
        sv[BN + n] = p.ch_shift ? __ldg(p.ch_shift + ng) : 0.f;
        sv[2 * BN + n] = p.act == E4S_ACT_PRELU ? __ldg(p.act_prelu + ng) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");           // the four epilogue warps only
      mbar_wait(bar_afull + 8 * set, use & 1);
      tc_fence_after();
      if (warp == 8 && lane == 0) hl_trace(1, it, 1);
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        if (ph >= P) break;
        const uint32_t tacc = tmem_base + (uint32_t)((set * P + ph) * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float acc[16];
          if (!(dbg & 2)) {
            tmem_ld16(tacc + (uint32_t)c0, acc);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
          }
          if (!(dbg & 1) && (rowlive & (1u << ph))) tc_epilogue16_sv(p, acc, jb.nt * BN + c0, c0, BN, sv, er[ph], dbg);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aempty + 8 * set);
      if (warp == 8 && lane == 0) hl_trace(1, it, 2);
    }
  } else if (warp == HL_MMA_WARP) {
    // =========================== MMA issuer (whole warp walks the loops, one elected lane issues) ==
    {
      int hg = 0, bc = 0;                                    // running halo-fill and weight-chunk counters
      for (int it = 0; it < my_jobs; ++it) {
        const int set = nsets == 2 ? (it & 1) : 0;
        const int use = nsets == 2 ? (it >> 1) : it;
        if (lane == 0) hl_trace(2, it, 0);
        mbar_wait(bar_aempty + 8 * set, (use & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) hl_trace(2, it, 1);
        for (int g = 0; g < G; ++g, ++hg) {
          const int hs = hg & 1;
          mbar_wait(bar_hfull + 8 * hs, (hg >> 1) & 1);
          if (lane == 0) hl_trace(2, it, 2);
          tc_fence_after();
          const uint32_t h_hi = smem_base + hs * HALO_BYTES;
          // descriptors are affine in the byte address: desc(a + off) = desc(a) + (off >> 4)  (14-bit field, smem < 256 KB)
          const uint64_t dah0 = umma_smem_desc_sbo(h_hi, HL_HP * 128), dal0 = umma_smem_desc_sbo(h_hi + HL_PLANE, HL_HP * 128);
          for (int pl = 0; pl < PL; ++pl) {
            const uint32_t tacc = tmem_base + (uint32_t)((set * P + pl * PM) * BN);
            for (int c = 0; c < CPG; ++c, ++bc) {
              const int bs = bc % BST;
              mbar_wait(bar_bfull + 8 * bs, (bc / BST) & 1);
              tc_fence_after();
              const uint32_t b_hi = smem_base + B_OFF + bs * STAGE_B;
              const uint64_t dbh0 = umma_smem_desc(b_hi), dbl0 = umma_smem_desc(b_hi + PM * B_BYTES);
              if (elect_one()) {
                for (int tt = 0; tt < ((g_hl_dbg & 4) ? 0 : tpc); ++tt) {
                  const int tap = c * tpc + tt;
                  if (tap >= 9) break;
                  const uint32_t aoff = ((uint32_t)((tap / 3) * HL_HP + (tap % 3)) * 128u) >> 4;
                  const uint32_t boff = (uint32_t)(tt * cin_eff * 2) >> 4;
                  for (int k = 0; k < ksteps; ++k) {
                    const uint64_t dah = dah0 + aoff + 2 * k, dal = dal0 + aoff + 2 * k;
                    const uint64_t dbh = dbh0 + boff + 2 * k, dbl = dbl0 + boff + 2 * k;
                    umma_bf16(tacc, dal, dbh, IDESC, (g | tap | k) != 0);
                    umma_bf16(tacc, dah, dbl, IDESC, 1);
                    umma_bf16(tacc, dah, dbh, IDESC, 1);
                  }
                }
                umma_commit(bar_bempty + 8 * bs);
              }
              __syncwarp();
            }
          }
          if (elect_one()) umma_commit(bar_hempty + 8 * hs);  // every tap of every phase has read this halo
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar_afull + 8 * set);
        __syncwarp();
        if (lane == 0) hl_trace(2, it, 3);
      }
    }
    __syncwarp();
  } else {
    // =========================== weight loader ====================================================
    {
      const int64_t tile_bytes = 2 * (int64_t)B_BYTES;
      int bc = 0;
      for (int it = 0; it < my_jobs; ++it) {
        const HlJob jb = decode(it);
        for (int g = 0; g < G; ++g)
          for (int pl = 0; pl < PL; ++pl)
            for (int c = 0; c < CPG; ++c, ++bc) {
              const int bs = bc % BST;
              const int kc = tpc == 1 ? g * 9 + c : c;       // packed chunk order: channel group outer, tap inner (pack_weights_tc)
              mbar_wait(bar_bempty + 8 * bs, ((bc / BST) & 1) ^ 1);
              if (g_hl_dbg & 64) {
                if (elect_one()) mbar_arrive(bar_bfull + 8 * bs);
              } else if (elect_one()) {
                const uint32_t dst = smem_base + B_OFF + bs * STAGE_B;
                mbar_arrive_expect_tx(bar_bfull + 8 * bs, STAGE_B);
                for (int q = 0; q < PM; ++q) {               // hi tiles of the merged phases back to back, then the lo tiles
                  const uint8_t* src = wpk + (((int64_t)(pl * PM + q) * n_tiles + jb.nt) * num_kc + kc) * tile_bytes;
                  bulk_g2s(dst + q * B_BYTES, src, B_BYTES, bar_bfull + 8 * bs);
                  bulk_g2s(dst + (PM + q) * B_BYTES, src + B_BYTES, B_BYTES, bar_bfull + 8 * bs);
                }
              }
              __syncwarp();
            }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == HL_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
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

static int g_halo_sm_count = 0;

// geometry the halo kernel takes (everything else stays on the gather kernel)
bool tc_halo_geometry_ok(const E4SConv* p) {
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  if (!up && !(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && p->in_shift == 0)) return false;
  if (p->hin % HL_TH || p->win % HL_TW) return false;
  if (!(p->cin == 32 || p->cin % 64 == 0)) return false;
  const int bn = tc_block_n(p->cout);
  if ((up ? 4 : 1) * bn > 512) return false;
  return true;
}

bool tc_halo_eligible(const E4SConv* p) {
  if (p->labels) return false;                               // per-pixel regions: region-job list (e4s_conv_tc_regions) or gather
  return tc_halo_geometry_ok(p);
}

template <int BN>
static int launch_halo(const E4SConv* p, const void* wpk, cudaStream_t s, const int4* rjobs = nullptr, const int* rjob_count = nullptr,
                       int rjob_host_count = 0) {
  static bool attr_set = false;
  const int pm = hl_phase_merge(BN, p->mode == E4S_CONV_UP2_POLYPHASE);
  const int smem_bytes = hl_smem_bytes(BN, pm);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_MAX);
    if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "conv_tc(halo): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (g_halo_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_halo_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_halo_sm_count <= 0) g_halo_sm_count = 148;
  }
  const int tiles_x = p->win / HL_TW, tiles_y = p->hin / HL_TH, n_tiles = p->cout / BN;
  const int64_t total = rjobs ? (int64_t)rjob_host_count * n_tiles : (int64_t)p->batch * tiles_x * tiles_y * n_tiles;
  E4S_REQUIRE(total > 0 && total < 0x7fffffff, "conv_tc(halo): bad job count");
  const unsigned grid = (unsigned)(total < g_halo_sm_count ? total : g_halo_sm_count);
  conv_tc_halo_kernel<BN><<<grid, HL_THREADS, smem_bytes, s>>>(*p, static_cast<const uint8_t*>(wpk), tiles_x, tiles_y, n_tiles,
                                                                     (int)total, rjobs, rjob_count);
  return check_launch("e4s_conv_tc(halo)");
}

int tc_halo_set_flags(int flags) {
  cudaError_t e = cudaMemcpyToSymbol(g_hl_dbg, &flags, sizeof(int));
  return e == cudaSuccess ? E4S_OK : fail(E4S_ERR_CUDA, "halo flags: %s", cudaGetErrorString(e));
}

int tc_halo_set_trace(void* buf, int cap_records) {
  unsigned long long* ptr = static_cast<unsigned long long*>(buf);
  int zero = 0;
  cudaError_t e = cudaMemcpyToSymbol(g_hl_trace, &ptr, sizeof(ptr));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_hl_trace_cap, &cap_records, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_hl_trace_n, &zero, sizeof(int));
  return e == cudaSuccess ? E4S_OK : fail(E4S_ERR_CUDA, "halo trace: %s", cudaGetErrorString(e));
}

int tc_launch_halo(const E4SConv* p, const void* wpk, cudaStream_t s, const int4* rjobs, const int* rjob_count, int rjob_host_count) {
  switch (tc_block_n(p->cout)) {
    case 256: return launch_halo<256>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    case 128: return launch_halo<128>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    case 64: return launch_halo<64>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
    case 32: return launch_halo<32>(p, wpk, s, rjobs, rjob_count, rjob_host_count);
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

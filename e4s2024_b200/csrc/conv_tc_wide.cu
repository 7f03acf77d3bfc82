// tcgen05 "wide" gather convolution: one im2col A tile feeds ALL 512 TMEM columns.
//
// The gather engine (conv_tc.cu) is bound by its A producers (global gather -> modulate -> bf16 hi/lo split -> swizzled
// st.shared), not by the tensor pipe: 58-62 % tensor-pipe utilisation at N = 256 and 35 % at N = 128
// (profiles/r1_ncu_full_v4_summary.tsv).  Its grid repeats that producer work once per 256 output columns and, for the
// poly-phase up-convolution, once per phase (blockIdx.z) although the four phases of an input pixel read the SAME 3x3
// input window.  Here one CTA owns a 16x8 pixel tile and accumulates four 128-column "slots" in TMEM from one A stream:
//
//   up-convolution (UP):  slot = phase (2x2 output pixels of the input pixel), 128 output channels per CTA
//   same resolution:      slot = 128-channel block, 512 output channels per CTA
//
// so the producer work per output element drops 4x (up, cout = 128), 2x (up, cout >= 256; same resolution, cout = 512).
// B (weights) streams through its own ring in half-steps of 256 rows (two slots) per A chunk.
//
// Regional modulation (reference model.py:395-398) multiplies the A row by the style of the OUTPUT pixel's region.  The
// four phases of an input pixel are four different output pixels: the A tile can be shared only if they lie in the same
// region.  That is decided per tile from the label map: tiles whose rows are phase-uniform take the merged path, a tile
// with any mixed row runs four passes (one per phase, N = 128, each modulated with that phase's regions) -- exact for
// every mask, fast where masks are piecewise constant at the layer's resolution.
//
// Warp roles (320 threads, one CTA per SM): warps 0-7 A producers then epilogue, warp 8 MMA issuer + TMEM owner,
// warp 9 weight loader (cp.async.bulk of the pre-swizzled tiles packed by e4s_pack_weights_tc).
#include "tc_ptx.cuh"

namespace e4s {

constexpr int WD_THREADS = 10 * 32;
constexpr int WD_A_STAGES = 2, WD_B_STAGES = 2;
constexpr int WD_A_STAGE = 2 * TC_A_BYTES;            // A hi | lo: 32 KB
constexpr int WD_SLOT_BYTES = 128 * 128;              // 128 weight rows of one 64-wide K chunk (hi or lo): 16 KB
constexpr int WD_B_STAGE = 4 * WD_SLOT_BYTES;         // two slots, hi | lo: 64 KB
constexpr int WD_SMEM = WD_A_STAGES * WD_A_STAGE + WD_B_STAGES * WD_B_STAGE + TC_BM * 16 + 256 + 1024;

struct WdRow {
  int b, y, x;        // sample and tile-grid coordinates (input pixel for UP, output pixel otherwise)
  uint32_t r4;        // region of the row's output pixel per slot, one byte each
};

// PAIR = true: launched as clusters of two CTAs (adjacent tiles, same channel block) that execute ONE cta_group::2 MMA
// (M = 256): each CTA gathers its own A tile and loads HALF of every B sub-step (its 128 of the 256 rows), so the weight
// bytes written to and read from each SM's shared memory halve.  The 1-CTA kernel needs ~147 B/cycle of shared-memory
// traffic at full MMA rate (A 4 KB + B 8 KB read per 128-cycle MMA, B 64 KB written per 1536 cycles) against 128
// available and sits at 62-71 % tensor-pipe-active; the pair needs ~94.  Rank 0 issues every MMA; rank 1's MMA warp
// relays "my A / B stage is full" to rank 0; stage releases and the accumulator barrier are multicast commits.
template <bool UP, bool PAIR>
__global__ void __launch_bounds__(WD_THREADS, 1) conv_tc_wide_kernel(const E4SConv p, const uint8_t* __restrict__ wpk, const int bn_packed,
                                                                     const int nt_packed) {
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;   // device-side launch predicate (e4s_b200.h)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int B_OFF = WD_A_STAGES * WD_A_STAGE;
  WdRow* rows = reinterpret_cast<WdRow*>(smem + B_OFF + WD_B_STAGES * WD_B_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF + WD_B_STAGES * WD_B_STAGE + TC_BM * 16);
  constexpr int NB = PAIR ? 4 : 2;                    // B ring stages (pair: half-size stages, twice as many)
  constexpr int BSTAGE = PAIR ? WD_B_STAGE / 2 : WD_B_STAGE;
  const uint32_t bar_afull = smem_u32(bars);          // 2
  const uint32_t bar_aempty = bar_afull + 16;         // 2
  const uint32_t bar_bfull = bar_aempty + 16;         // up to 4
  const uint32_t bar_bempty = bar_bfull + 32;         // up to 4
  const uint32_t bar_acc = bar_bempty + 32;
  const uint32_t bar_pafull = bar_acc + 8;            // 2: peer's A stage is full (pair, rank 0 only)
  const uint32_t bar_pbfull = bar_pafull + 16;        // 4: peer's B stage is full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  uint32_t* pair_flag = tmem_slot + 1;                // this CTA's "tile has mixed rows" flag, read by the peer
  const uint32_t crank = PAIR ? cluster_rank() : 0u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int num_kc = 9 * p.cin / 64;                  // cin % 64 == 0: chunk = (64-channel group, tap), group outer
  const int gw = UP ? p.win : p.wout, gh = UP ? p.hin : p.hout;

  // ---- one-time setup ---------------------------------------------------------------------------
  int mixed_row = 0;
  if (tid < TC_BM) {
    const int tx_n = gw >> 3, ty_n = gh >> 4;
    int t = (int)blockIdx.x;
    const int txi = t % tx_n;
    t /= tx_n;
    const int tyi = t % ty_n;
    WdRow rw;
    rw.b = t / ty_n;
    rw.y = tyi * 16 + (tid >> 3);
    rw.x = txi * 8 + (tid & 7);
    rw.r4 = 0;
    if (p.labels) {
      if (UP) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int sy = nearest_src(2 * rw.y + (s >> 1), p.lab_h, p.hout), sx = nearest_src(2 * rw.x + (s & 1), p.lab_w, p.wout);
          rw.r4 |= (uint32_t)p.labels[((int64_t)rw.b * p.lab_h + sy) * p.lab_w + sx] << (8 * s);
        }
        mixed_row = rw.r4 != (rw.r4 & 0xffu) * 0x01010101u;
      } else {
        const int sy = nearest_src(rw.y, p.lab_h, p.hout), sx = nearest_src(rw.x, p.lab_w, p.wout);
        rw.r4 = (uint32_t)p.labels[((int64_t)rw.b * p.lab_h + sy) * p.lab_w + sx] * 0x01010101u;
      }
    }
    rows[tid] = rw;
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, TC_PRODUCER_WARPS);
      mbar_init(bar_aempty + 8 * s, 1);
      mbar_init(bar_pafull + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
      mbar_init(bar_pbfull + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == TC_PRODUCER_WARPS) {
    if (PAIR) tmem_alloc2(smem_u32(tmem_slot), 512);
    else tmem_alloc(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  int mixed = __syncthreads_or(mixed_row);             // any row whose four phases straddle a region boundary
  if (PAIR) {                                          // both CTAs of a pair must walk the same step sequence
    if (tid == 0) *pair_flag = (uint32_t)mixed;
    cluster_sync_all();                                // also: barrier inits visible to the peer before any remote arrive / multicast commit
    mixed |= (int)ld_shared_cluster_u32(mapa_u32(smem_u32(pair_flag), crank ^ 1u));
  }
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const int npass = mixed ? 4 : 1;                     // mixed tile: one pass per phase, N = 128
  const int nsub = mixed ? 1 : 2;                      // B half-steps per A chunk
  const int nsteps = npass * num_kc;

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== A producers =====================================================
    const int cg = tid & 7;           // 8-channel group inside the 64-wide K chunk
    const int r0 = tid >> 3;          // rows r0, r0+32, r0+64, r0+96
    int rb[4], ry[4], rx[4];
    uint32_t rr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const WdRow rw = rows[r0 + 32 * i];
      rb[i] = rw.b;
      const int st = UP ? 1 : p.stride;                               // same-resolution mode also takes the stride-2 3x3 convs (encoder)
      ry[i] = rw.y * st - 1;          // input row of tap ky = 0 (3x3, pad 1; the poly-phase taps read the same window)
      rx[i] = rw.x * st - 1;
      rr[i] = rw.r4;
    }
    float4 sreg[4][2];                // modulation of this thread's 8 channels per row: changes every 9 chunks (and per pass)
    int cached_g = -1;
    float4 v[4][2];
    bool ok[4];
    auto prefetch = [&](int step) {
      const int kc = step % num_kc;
      const int g = kc / 9, tap = kc - g * 9;
      const int ci = g * 64 + cg * 8;
      const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int iy = ry[i] + ky, ix = rx[i] + kx;
        ok[i] = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win;
        if (ok[i]) {
          const float4* src = reinterpret_cast<const float4*>(p.x + (((int64_t)rb[i] * p.hin + iy) * p.win + ix) * p.x_pitch + ci);
          v[i][0] = __ldg(src);          // allocate in L1: the 9 taps re-read these lines
          v[i][1] = __ldg(src + 1);
        }
      }
    };

    prefetch(0);
    for (int step = 0; step < nsteps; ++step) {
      const int s = step & 1;
      const int pass = step / num_kc, kc = step - pass * num_kc;
      mbar_wait(bar_aempty + 8 * s, ((step >> 1) & 1) ^ 1);
      uint8_t* a_hi = smem + s * WD_A_STAGE;
      uint8_t* a_lo = a_hi + TC_A_BYTES;
      const int g = kc / 9;
      if (p.smod && pass * 1024 + g != cached_g) {
        cached_g = pass * 1024 + g;
        const int ci = g * 64 + cg * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int reg = (rr[i] >> (8 * pass)) & 0xff;      // merged tiles have one region per row: pass 0 == every phase
          const float4* sp = reinterpret_cast<const float4*>(p.smod + ((int64_t)rb[i] * p.regions + reg) * p.cin + ci);
          sreg[i][0] = __ldg(sp);
          sreg[i][1] = __ldg(sp + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = r0 + 32 * i;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ok[i]) {
          f[0] = v[i][0].x; f[1] = v[i][0].y; f[2] = v[i][0].z; f[3] = v[i][0].w;
          f[4] = v[i][1].x; f[5] = v[i][1].y; f[6] = v[i][1].z; f[7] = v[i][1].w;
          if (p.smod) {
            const float4 s0 = sreg[i][0], s1 = sreg[i][1];
            f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
            f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
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
        const uint32_t off = row * 128 + ((cg ^ (row & 7)) << 4);   // 128B swizzle: 16B chunk index ^= row % 8
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      // publish the stage BEFORE issuing the next chunk's global loads (fence.proxy.async waits for outstanding loads)
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull + 8 * s);
      if (step + 1 < nsteps) prefetch(step + 1);
    }

    // =========================== epilogue (fast form only, see tc_wide_eligible) ===================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int hh = warp >> 2;               // slot pair (2hh, 2hh + 1)
    const WdRow rw = rows[q * 32 + lane];
    const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const float corr = tc_acc_unbias(p, num_kc * 4 * 3);                 // accumulate steps per output (tc_ptx.cuh)
#pragma unroll 1
    for (int sl = 2 * hh; sl < 2 * hh + 2; ++sl) {
      const int oy = UP ? 2 * rw.y + (sl >> 1) : rw.y, ox = UP ? 2 * rw.x + (sl & 1) : rw.x;
      const int reg = (rw.r4 >> (8 * sl)) & 0xff;
      const int n_base = UP ? (int)blockIdx.y * 128 : (int)blockIdx.y * 512 + sl * 128;
      const int64_t pix = ((int64_t)rw.b * p.hout + oy) * p.wout + ox;
      const float* drow = p.demod ? p.demod + ((int64_t)rw.b * p.regions + reg) * p.cout : nullptr;
      const float nz = p.noise ? nw * __ldg(p.noise + (int64_t)rw.b * p.noise_sb + (int64_t)oy * p.wout + ox) : 0.f;
      float* optr = p.out + pix * p.out_pitch + n_base;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        float acc[16];
        float4 mul[4], add[4], slp[4];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int n = n_base + c0 + 4 * qd;
          mul[qd] = drow ? ldg4(drow + n) : make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.ch_scale) {
            const float4 sc = ldg4(p.ch_scale + n);
            mul[qd].x *= sc.x; mul[qd].y *= sc.y; mul[qd].z *= sc.z; mul[qd].w *= sc.w;
          }
          mul[qd].x *= corr; mul[qd].y *= corr; mul[qd].z *= corr; mul[qd].w *= corr;
          add[qd] = p.ch_shift ? ldg4(p.ch_shift + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          slp[qd] = make_float4(tc_epi_slope(p, n), tc_epi_slope(p, n + 1), tc_epi_slope(p, n + 2), tc_epi_slope(p, n + 3));
        }
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * 128 + c0), acc);   // warp-collective
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          float4 a;
          a.x = fmaf(acc[4 * qd], mul[qd].x, add[qd].x + nz); a.y = fmaf(acc[4 * qd + 1], mul[qd].y, add[qd].y + nz);
          a.z = fmaf(acc[4 * qd + 2], mul[qd].z, add[qd].z + nz); a.w = fmaf(acc[4 * qd + 3], mul[qd].w, add[qd].w + nz);
          a.x = (a.x < 0.f ? a.x * slp[qd].x : a.x) * gain; a.y = (a.y < 0.f ? a.y * slp[qd].y : a.y) * gain;
          a.z = (a.z < 0.f ? a.z * slp[qd].z : a.z) * gain; a.w = (a.w < 0.f ? a.w * slp[qd].w : a.w) * gain;
          reinterpret_cast<float4*>(optr + c0)[qd] = a;
        }
      }
    }
    tc_fence_before();
  } else if (warp == TC_PRODUCER_WARPS) {
    // =========================== MMA issuer ========================================================
    constexpr uint32_t D_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);    // SBO 1024, version 1, SWIZZLE_128B
    const int nmma = mixed ? 128 : 256;                                            // N of one MMA
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    // B_lo follows this CTA's rows of B_hi (pair: half of the MMA's N rows live in each CTA)
    const uint32_t lo_off = (uint32_t)((nmma / (PAIR ? 2 : 1)) * 128) >> 4;
    int bcount = 0;
    if (PAIR && crank != 0) {
      // rank 1: relay "stage full" to rank 0 (its producers / loader only signal their own CTA's barriers)
      if (elect_one()) {
        for (int step = 0; step < nsteps; ++step) {
          const int s = step & 1;
          mbar_wait(bar_afull + 8 * s, (step >> 1) & 1);
          mbar_arrive_remote(mapa_u32(bar_pafull + 8 * s, 0));
          for (int sub = 0; sub < nsub; ++sub, ++bcount) {
            const int bs = bcount % NB;
            mbar_wait(bar_bfull + 8 * bs, (bcount / NB) & 1);
            mbar_arrive_remote(mapa_u32(bar_pbfull + 8 * bs, 0));
          }
        }
      }
      __syncwarp();
    } else {
      for (int step = 0; step < nsteps; ++step) {
        const int s = step & 1;
        const int pass = step / num_kc, kc = step - pass * num_kc;
        mbar_wait(bar_afull + 8 * s, (step >> 1) & 1);
        if (PAIR) mbar_wait_cluster(bar_pafull + 8 * s, (step >> 1) & 1);
        tc_fence_after();
        const uint32_t a_h = (((smem_base + s * WD_A_STAGE) >> 4) & 0x3FFFu) | (1u << 16), a_l = a_h + (TC_A_BYTES >> 4);
        for (int sub = 0; sub < nsub; ++sub, ++bcount) {
          const int bs = bcount % NB;
          mbar_wait(bar_bfull + 8 * bs, (bcount / NB) & 1);
          if (PAIR) mbar_wait_cluster(bar_pbfull + 8 * bs, (bcount / NB) & 1);
          tc_fence_after();
          const uint32_t b_h = (((smem_base + B_OFF + bs * BSTAGE) >> 4) & 0x3FFFu) | (1u << 16), b_l = b_h + lo_off;
          const uint32_t tacc = tmem_acc + (uint32_t)(mixed ? pass * 128 : sub * 256);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t first = (uint32_t)((kc | k) != 0);
              if (PAIR) {
                asm volatile(
                    "{\n\t"
                    ".reg .pred p, t;\n\t"
                    ".reg .b64 dah, dal, dbh, dbl;\n\t"
                    "setp.ne.b32 p, %6, 0;\n\t"
                    "setp.eq.b32 t, 0, 0;\n\t"
                    "mov.b64 dah, {%1, %5};\n\t"
                    "mov.b64 dal, {%2, %5};\n\t"
                    "mov.b64 dbh, {%3, %5};\n\t"
                    "mov.b64 dbl, {%4, %5};\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], dal, dbh, %7, p;\n\t"      // small terms first
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbl, %7, t;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbh, %7, t;\n\t"
                    "}" ::"r"(tacc),
                    "r"(a_h + 2 * k), "r"(a_l + 2 * k), "r"(b_h + 2 * k), "r"(b_l + 2 * k), "r"(D_HI), "r"(first), "r"(idesc)
                    : "memory");
              } else {
                asm volatile(
                    "{\n\t"
                    ".reg .pred p, t;\n\t"
                    ".reg .b64 dah, dal, dbh, dbl;\n\t"
                    "setp.ne.b32 p, %6, 0;\n\t"
                    "setp.eq.b32 t, 0, 0;\n\t"
                    "mov.b64 dah, {%1, %5};\n\t"
                    "mov.b64 dal, {%2, %5};\n\t"
                    "mov.b64 dbh, {%3, %5};\n\t"
                    "mov.b64 dbl, {%4, %5};\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %7, p;\n\t"      // small terms first
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %7, t;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %7, t;\n\t"
                    "}" ::"r"(tacc),
                    "r"(a_h + 2 * k), "r"(a_l + 2 * k), "r"(b_h + 2 * k), "r"(b_l + 2 * k), "r"(D_HI), "r"(first), "r"(idesc)
                    : "memory");
              }
            }
            if (PAIR) umma_commit_mc2(bar_bempty + 8 * bs, 3);
            else umma_commit(bar_bempty + 8 * bs);
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (PAIR) umma_commit_mc2(bar_aempty + 8 * s, 3);
          else umma_commit(bar_aempty + 8 * s);
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (PAIR) umma_commit_mc2(bar_acc, 3);
        else umma_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // =========================== weight loader =====================================================
    // packed chunk (e4s_pack_weights_tc): per (n_tile, kc): [hi tiles of the P phases | lo tiles of the P phases], bn_packed rows each
    const int64_t t_bytes = (int64_t)bn_packed * 128;
    const int phases = UP ? 4 : 1;
    int bcount = 0;
    for (int step = 0; step < nsteps; ++step) {
      const int pass = step / num_kc, kc = step - pass * num_kc;
      for (int sub = 0; sub < nsub; ++sub, ++bcount) {
        const int bs = bcount % NB;
        mbar_wait(bar_bempty + 8 * bs, ((bcount / NB) & 1) ^ 1);
        const uint32_t dst = smem_base + B_OFF + bs * BSTAGE;
        if (elect_one()) {
          if (UP) {
            const int cb = (int)blockIdx.y;                            // 128-channel block of this CTA (pair)
            const int nt = cb * 128 / bn_packed;
            const uint8_t* base = wpk + ((int64_t)nt * num_kc + kc) * (2 * phases * t_bytes) + (int64_t)(cb * 128 % bn_packed) * 128;
            if (PAIR) {
              if (mixed) {                                             // phase `pass`, this CTA's 64 of the 128 channels: hi | lo
                const int64_t roff = (int64_t)crank * 64 * 128;
                mbar_arrive_expect_tx(bar_bfull + 8 * bs, WD_SLOT_BYTES);
                bulk_g2s(dst, base + pass * t_bytes + roff, WD_SLOT_BYTES / 2, bar_bfull + 8 * bs);
                bulk_g2s(dst + WD_SLOT_BYTES / 2, base + (phases + pass) * t_bytes + roff, WD_SLOT_BYTES / 2, bar_bfull + 8 * bs);
              } else {                                                 // this CTA's phase of the pair (2sub + rank): hi | lo
                const int ph = 2 * sub + (int)crank;
                mbar_arrive_expect_tx(bar_bfull + 8 * bs, 2 * WD_SLOT_BYTES);
                bulk_g2s(dst, base + ph * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
                bulk_g2s(dst + WD_SLOT_BYTES, base + (phases + ph) * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
              }
            } else if (mixed) {                                        // one phase: hi 128 rows | lo 128 rows
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, 2 * WD_SLOT_BYTES);
              bulk_g2s(dst, base + pass * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
              bulk_g2s(dst + WD_SLOT_BYTES, base + (phases + pass) * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
            } else {                                                   // phases 2sub, 2sub+1: hi hi | lo lo
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, 4 * WD_SLOT_BYTES);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                bulk_g2s(dst + j * WD_SLOT_BYTES, base + (2 * sub + j) * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
                bulk_g2s(dst + (2 + j) * WD_SLOT_BYTES, base + (phases + 2 * sub + j) * t_bytes, WD_SLOT_BYTES, bar_bfull + 8 * bs);
              }
            }
          } else {                                                     // channels [512*by + 256*sub, +256): one packed 256-row chunk (hi | lo)
            const int nt = (int)blockIdx.y * 2 + sub;
            const uint8_t* src = wpk + ((int64_t)nt * num_kc + kc) * (2 * t_bytes);
            if (PAIR) {                                                // this CTA's 128 of the 256 rows: hi | lo
              const int64_t roff = (int64_t)crank * WD_SLOT_BYTES;
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, 2 * WD_SLOT_BYTES);
              bulk_g2s(dst, src + roff, WD_SLOT_BYTES, bar_bfull + 8 * bs);
              bulk_g2s(dst + WD_SLOT_BYTES, src + t_bytes + roff, WD_SLOT_BYTES, bar_bfull + 8 * bs);
            } else {
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, 4 * WD_SLOT_BYTES);
              bulk_g2s(dst, src, 4 * WD_SLOT_BYTES, bar_bfull + 8 * bs);
            }
          }
        }
        __syncwarp();
      }
    }
  }

  __syncthreads();
  if (PAIR) cluster_sync_all();                        // the peer may still be reading this CTA's operands / accumulating into its TMEM
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_acc, 512);
    else tmem_dealloc(tmem_acc, 512);
  }
}

// What the wide kernel takes: 3x3 stride-1 (or poly-phase up) layers on full 16x8 tiles with 64-channel K groups, the
// branch-free epilogue, and enough output columns to fill TMEM (up: 128-channel blocks x 4 phases; same: 512 channels).
bool tc_wide_eligible(const E4SConv* p) {
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  if (!tc_epi_is_fast(*p) || p->in_mean || p->in_shift || p->in_square || p->rgb || !p->out) return false;
  if (!(p->kh == 3 && p->kw == 3)) return false;
  if (!up && !((p->stride == 1 || p->stride == 2) && p->pad == 1)) return false;
  if (p->cin % 64) return false;
  const int gh = up ? p->hin : p->hout, gw = up ? p->win : p->wout;
  if (gh % 16 || gw % 8) return false;
  if (p->regions > 255) return false;
  // stride 2: only with enough tiles to fill the machine (512 -> 512 @64^2 -> 32^2, B = 16: 0.35 -> 0.17 ms); a 32-tile launch is faster on
  // the gather kernel, whose grid also splits the output channels
  if (!up && p->stride == 2 && (int64_t)p->batch * (gh / 16) * (gw / 8) < 100) return false;
  return up ? (p->cout % 128 == 0) : (p->cout % 512 == 0);
}

// -1: read E4S_TC_WIDE_PAIR.  Default OFF: the cta_group::2 variant is correct (same tests) but measured 2-7 % SLOWER than
// the 1-CTA kernel on every wide layer (profiles/r1_layers_v12_pair.jsonl) -- the 1-CTA kernel already runs at the
// sustained (power-limited) bf16 rate cuBLAS reaches on this part, so halving the per-SM weight traffic buys nothing.
static int g_wide_pair = -1;

template <bool UP, bool PAIR>
static int launch_wide_t(const E4SConv* p, const void* wpk, cudaStream_t s) {
  static bool attr_set_dev[E4S_MAX_DEVICES] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wide_kernel<UP, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM);
    if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "conv_tc(wide): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int gh = UP ? p->hin : p->hout, gw = UP ? p->win : p->wout;
  const int64_t tiles = (int64_t)p->batch * (gh / 16) * (gw / 8);
  E4S_REQUIRE(tiles > 0 && tiles < 0x7fffffff, "conv_tc(wide): bad tile count");
  const int bn = tc_block_n(p->cout);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)tiles, (unsigned)(UP ? p->cout / 128 : p->cout / 512), 1);
  cfg.blockDim = dim3(WD_THREADS);
  cfg.dynamicSmemBytes = WD_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t* wp = static_cast<const uint8_t*>(wpk);
  const int ntp = p->cout / bn;
  cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc_wide_kernel<UP, PAIR>, *p, wp, bn, ntp);
  if (le != cudaSuccess) return fail(E4S_ERR_CUDA, "e4s_conv_tc(wide): launch: %s", cudaGetErrorString(le));
  return check_launch("e4s_conv_tc(wide)");
}

int tc_launch_wide(const E4SConv* p, const void* wpk, cudaStream_t s) {
  if (g_wide_pair < 0) {
    const char* e = getenv("E4S_TC_WIDE_PAIR");
    g_wide_pair = (e && e[0] == '1') ? 1 : 0;
  }
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  const int gh = up ? p->hin : p->hout, gw = up ? p->win : p->wout;
  const bool pair = g_wide_pair && (((int64_t)p->batch * (gh / 16) * (gw / 8)) % 2 == 0);    // pairs = adjacent tiles along grid.x
  if (up) return pair ? launch_wide_t<true, true>(p, wpk, s) : launch_wide_t<true, false>(p, wpk, s);
  return pair ? launch_wide_t<false, true>(p, wpk, s) : launch_wide_t<false, false>(p, wpk, s);
}

}  // namespace e4s

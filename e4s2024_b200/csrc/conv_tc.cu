// tcgen05 (5th-generation tensor core) implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels, BN channels] (fp32, TMEM)  +=  A[128, 64] * B[BN, 64]^T    per K chunk of 64
//
// fp32-class accuracy on bf16 tensor cores: every operand is split v = hi + lo (two bf16), and each
// product is issued as three MMAs  hi*hi + hi*lo + lo*hi  accumulating in fp32 (error ~2^-16 relative,
// measured 6e-5 on the 256^2 generator).
//
// Warp roles (320 threads, one CTA per SM):
//   warps 0-7  A producers: gather the im2col rows of this K chunk straight from the NHWC fp32 activations
//              (zero padding, stride, nearest-upsampled input view, poly-phase up-conv taps), apply
//              InstanceNorm and the per-pixel REGION style modulation  x * s[b, r(p), ci]  (each GEMM row
//              carries its own region -> one pass over the tile whatever the mask looks like), split to
//              bf16 hi/lo and store into the 128B-swizzled K-major UMMA layout.  After the main loop the
//              same warps run the fused epilogue: tcgen05.ld the accumulator, demodulate per (sample, region,
//              channel), add noise / bias / residual, activate, store NHWC fp32.
//   warp 8     MMA issuer (one elected lane): 12 tcgen05.mma (3 split terms x 4 K-steps of 16) per chunk;
//              tcgen05.commit releases the smem stage / signals the epilogue.  Owns the TMEM allocation.
//   warp 9     B loader: the weights were packed once (e4s_pack_weights_tc) into the exact swizzled smem image,
//              so each stage is two cp.async.bulk (TMA bulk-copy engine) transfers completing on an mbarrier.
#include "tc_ptx.cuh"

namespace e4s {

struct TcRow {
  int b, oy, ox, r;
};

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// SPLITK is a template parameter: the plain instantiation sits at the register cap (168) and must not pay for the extra state.
template <int BN, bool SPLITK = false>
// ksplit > 1 (split-K, same-resolution / strided layers only): blockIdx.z selects a slice of num_kc / ksplit K chunks and the CTA writes its
// RAW partial accumulator to part[z][output pixel][cout]; e4s_conv_tc_splitk's second kernel adds the slices in index order and applies the
// epilogue.  For the generator's 4^2-16^2 layers: a handful of CTAs walking 72 chunks each otherwise (0.17-0.2 ms of pure latency).
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const E4SConv p, const uint8_t* __restrict__ wpk, const int64_t m_total,
                                                                  const int tile2d, const int ksplit, float* __restrict__ part) {
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;   // device-side launch predicate (e4s_b200.h)
  constexpr int STAGES = tc_stages(BN);
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = tc_stage_bytes(BN);
  const uint32_t IDESC = umma_idesc_fmt(umma_idesc(BN), p.tc_fmt);
  const bool f16 = p.tc_fmt == E4S_TC_F16;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // [stages x (A_hi | A_lo | B_hi | B_lo)] [rows: 128 x int4] [barriers]
  TcRow* rows = reinterpret_cast<TcRow*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + TC_BM * 16);
  const uint32_t bar_full = smem_u32(bars);                    // STAGES barriers
  const uint32_t bar_empty = bar_full + 8 * STAGES;            // STAGES barriers
  const uint32_t bar_acc = bar_empty + 8 * STAGES;             // accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* sv = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + TC_BM * 16 + 256);   // [mul | add | slope] x BN: this CTA's channels

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int phase_id = SPLITK ? 0 : (int)blockIdx.z, py = phase_id >> 1, px = phase_id & 1;
  const int n_tile = blockIdx.y;
  const int K = p.kh * p.kw * p.cin;
  const int num_kc_all = (K + TC_BK - 1) / TC_BK;   // K is zero-padded to a multiple of 64 in the packed weights
  const int num_kc = SPLITK ? num_kc_all / ksplit : num_kc_all;   // chunks of this CTA's K slice (the host makes ksplit divide the chunk count)
  const int kc_base = SPLITK ? (int)blockIdx.z * num_kc : 0;
  const bool up = p.mode == E4S_CONV_UP2_POLYPHASE;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid < TC_BM) {
    int64_t m = (int64_t)blockIdx.x * TC_BM + tid;
    TcRow rw{-1, 0, 0, 0};
    if (tile2d) {
      // 16 x 8 pixel tiles: the 3x3 halo of a tile is 18 x 10 pixels instead of 3 x 130 for a 1 x 128 row segment, so the
      // 9 taps of one 64-channel chunk (46 KB) are re-read from L1 instead of L2
      const int gw = up ? p.win : p.wout, gh = up ? p.hin : p.hout;
      const int tx_n = gw >> 3, ty_n = gh >> 4;
      int t = (int)blockIdx.x;
      const int txi = t % tx_n;
      t /= tx_n;
      const int tyi = t % ty_n;
      rw.b = t / ty_n;
      const int yy = tyi * 16 + (tid >> 3), xx = txi * 8 + (tid & 7);
      rw.oy = up ? 2 * yy + py : yy;
      rw.ox = up ? 2 * xx + px : xx;
      if (p.labels) {
        int sy = nearest_src(rw.oy, p.lab_h, p.hout), sx = nearest_src(rw.ox, p.lab_w, p.wout);
        rw.r = p.labels[((int64_t)rw.b * p.lab_h + sy) * p.lab_w + sx];
      }
    } else if (m < m_total) {
      if (up) {
        int hw = p.hin * p.win;
        rw.b = (int)(m / hw);
        int rem = (int)(m - (int64_t)rw.b * hw);
        int a = rem / p.win;
        rw.oy = 2 * a + py;
        rw.ox = 2 * (rem - a * p.win) + px;
      } else {
        int hw = p.hout * p.wout;
        rw.b = (int)(m / hw);
        int rem = (int)(m - (int64_t)rw.b * hw);
        rw.oy = rem / p.wout;
        rw.ox = rem - rw.oy * p.wout;
      }
      if (p.labels) {
        int sy = nearest_src(rw.oy, p.lab_h, p.hout), sx = nearest_src(rw.ox, p.lab_w, p.wout);
        rw.r = p.labels[((int64_t)rw.b * p.lab_h + sy) * p.lab_w + sx];
      }
    }
    rows[tid] = rw;
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, TC_PRODUCER_WARPS / 2 + 1);  // the 4 producer warps of the chunk's group + the B loader's expect_tx arrive
      mbar_init(bar_empty + 8 * s, 1);                     // one tcgen05.commit
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == TC_PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), 2 * BN);   // [main | small-term accumulator of the fp16 mode]
  {
    // the layer-wide epilogue vectors of this CTA's channels, fetched once while the pipeline fills (read per 16-column block from
    // global memory they put an L2 round trip on the critical path of every block of the epilogue)
    const float corr0 = tc_acc_unbias(p, num_kc * (TC_BK / 16) * (f16 ? 1 : 3));
    for (int n = tid; n < BN; n += TC_THREADS) {
      const int gn = n_tile * BN + n;
      sv[n] = (p.ch_scale ? __ldg(p.ch_scale + gn) : 1.f) * corr0;
      sv[BN + n] = p.ch_shift ? __ldg(p.ch_shift + gn) : 0.f;
      sv[2 * BN + n] = tc_epi_slope(p, gn);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== A producers =====================================================
    // Two producer groups (warps 0-3 / 4-7) take the even / odd K chunks: a group issues the loads of its NEXT chunk right after
    // publishing the current one, so two chunks of global loads are in flight per SM while each thread's fence.proxy.async (which
    // waits for the thread's own outstanding loads) only ever sees loads it is about to consume anyway.  (Keeping two chunks in
    // flight per THREAD does not work: the fence of chunk kc then waits for the loads of chunk kc+1.)
    constexpr int NR = 8;             // rows per thread: r0 + 16 i
    const int grp = warp >> 2;
    const int cg = tid & 7;           // 8-channel group inside the 64-wide K chunk
    const int r0 = (tid & 127) >> 3;  // rows r0, r0+16, ..., r0+112
    int rb[NR], ry[NR], rx[NR];
    const float* rs[NR];              // modulation row (per pixel region)
    const float* rmean[NR];
    const float* rrstd[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const TcRow rw = rows[r0 + 16 * i];
      rb[i] = rw.b;
      if (up) {
        ry[i] = (rw.oy >> 1) - 1;     // input row of tap u = 0
        rx[i] = (rw.ox >> 1) - 1;
      } else {
        ry[i] = rw.oy * p.stride - p.pad;
        rx[i] = rw.ox * p.stride - p.pad;
      }
      const int bb = rw.b < 0 ? 0 : rw.b;
      rs[i] = p.smod ? p.smod + ((int64_t)bb * p.regions + rw.r) * p.cin : nullptr;
      rmean[i] = p.in_mean ? p.in_mean + (int64_t)bb * p.cin : nullptr;
      rrstd[i] = p.in_mean ? p.in_rstd + (int64_t)bb * p.cin : nullptr;
    }
    const int hv = p.hin << p.in_shift, wv = p.win << p.in_shift;
    const int kwid = up ? 3 : p.kw;
    const int ntaps = p.kh * p.kw;
    const bool grouped = (p.cin % 64) == 0;

    float4 v[NR][2];
    uint32_t okm = 0;
    auto prefetch = [&](int kc) {
      // K order of the packed weights (tc_chunk_k): 64-channel group outer, tap inner when cin % 64 == 0, so the taps of
      // one channel group run back to back over the same (L1-resident) input window; plain k = tap*cin + ci otherwise
      int tap, ci;
      bool tap_ok;
      if (grouped) {
        const int g = kc / ntaps;
        tap = kc - g * ntaps;
        ci = g * 64 + cg * 8;
        tap_ok = true;
      } else {
        const int k0 = kc * TC_BK + cg * 8;
        tap = k0 / p.cin;
        ci = k0 - tap * p.cin;
        tap_ok = k0 < K;
      }
      const int ky = tap / kwid, kx = tap - ky * kwid;
      okm = 0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        int iy = ry[i] + ky, ix = rx[i] + kx;
        const bool ok = tap_ok && rb[i] >= 0 && iy >= 0 && iy < hv && ix >= 0 && ix < wv;
        if (ok) {
          iy >>= p.in_shift;
          ix >>= p.in_shift;
          const float4* src = reinterpret_cast<const float4*>(p.x + (((int64_t)rb[i] * p.hin + iy) * p.win + ix) * p.x_pitch + ci);
          v[i][0] = __ldg(src);          // allocate in L1: the 9 taps re-read these lines
          v[i][1] = __ldg(src + 1);
          okm |= 1u << i;
        }
      }
    };

    if (grp < num_kc) prefetch(kc_base + grp);
    for (int kc = grp; kc < num_kc; kc += 2) {
      const int s = kc % STAGES;
      const uint32_t par = (kc / STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, par ^ 1);
      uint8_t* a_hi = smem + s * STAGE_BYTES;
      uint8_t* a_lo = a_hi + TC_A_BYTES;
      const int kca = kc_base + kc;                 // absolute chunk index
      const int ci = grouped ? (kca / ntaps) * 64 + cg * 8 : (kca * TC_BK + cg * 8) % p.cin;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int row = r0 + 16 * i;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if ((okm >> i) & 1u) {
          f[0] = v[i][0].x; f[1] = v[i][0].y; f[2] = v[i][0].z; f[3] = v[i][0].w;
          f[4] = v[i][1].x; f[5] = v[i][1].y; f[6] = v[i][1].z; f[7] = v[i][1].w;
          if (rmean[i]) {
            const float4 m0 = __ldg(reinterpret_cast<const float4*>(rmean[i] + ci)), m1 = __ldg(reinterpret_cast<const float4*>(rmean[i] + ci) + 1);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(rrstd[i] + ci)), q1 = __ldg(reinterpret_cast<const float4*>(rrstd[i] + ci) + 1);
            f[0] = (f[0] - m0.x) * q0.x; f[1] = (f[1] - m0.y) * q0.y; f[2] = (f[2] - m0.z) * q0.z; f[3] = (f[3] - m0.w) * q0.w;
            f[4] = (f[4] - m1.x) * q1.x; f[5] = (f[5] - m1.y) * q1.y; f[6] = (f[6] - m1.z) * q1.z; f[7] = (f[7] - m1.w) * q1.w;
          }
          if (rs[i]) {          // per-row style modulation (L1-resident rows of the table)
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(rs[i] + ci)), s1 = __ldg(reinterpret_cast<const float4*>(rs[i] + ci) + 1);
            f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
            f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
          }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tc_split2(f16, f[2 * j], f[2 * j + 1], hi[j], lo[j]);
        const uint32_t off = row * 128 + ((cg ^ (row & 7)) << 4);   // 128B swizzle: 16B chunk index ^= row % 8
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      // publish the stage BEFORE issuing the next chunk's global loads: fence.proxy.async waits for the thread's
      // outstanding memory operations, so a prefetch issued ahead of it is not a prefetch at all
      fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      if (kc + 2 < num_kc) prefetch(kc_base + kc + 2);
    }

    // =========================== epilogue ========================================================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = warp >> 2;             // column half
    const TcRow rw = rows[q * 32 + lane];
    const bool live = rw.b >= 0;
    const int64_t pix = live ? ((int64_t)rw.b * p.hout + rw.oy) * p.wout + rw.ox : 0;
    const int n_base = n_tile * BN + half * (BN / 2);
    const float* drow = (p.demod && live) ? p.demod + ((int64_t)rw.b * p.regions + rw.r) * p.cout : nullptr;
    float pw = 1.f;
    if (p.pixw && live) {
      int sy = nearest_src(rw.oy, p.lab_h, p.hout), sx = nearest_src(rw.ox, p.lab_w, p.wout);
      pw = __ldg(p.pixw + (int64_t)rw.b * p.pixw_sb + (int64_t)sy * p.lab_w + sx);
    }
    const float corr = tc_acc_unbias(p, num_kc * (TC_BK / 16) * (f16 ? 1 : 3));   // accumulate steps into the main accumulator      // 3 accumulating MMAs per K step of 16
    TcEpiRow er;
    er.pix = pix;
    er.drow = drow;
    er.pw = pw;
    er.nw = p.noise ? __ldg(p.noise_w) : 0.f;
    er.nrow = (p.noise && live) ? p.noise + (int64_t)rw.b * p.noise_sb + (int64_t)rw.oy * p.wout + rw.ox : nullptr;
    er.nz = (er.nrow && p.noise_sc == 0) ? er.nw * __ldg(er.nrow) : 0.f;
    if (SPLITK) {
      // split-K: the raw accumulator of this K slice -> part[z][pixel][cout] (the reduction kernel applies un-bias, demodulation, epilogue)
      float* prow = part + ((int64_t)blockIdx.z * m_total + pix) * p.cout + n_base;
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 16) {
        float acc[16];
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + c0), acc);   // warp-collective
        if (f16) {
          float acc2[16];
          tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + half * (BN / 2) + c0), acc2);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc2[j], 1.f / TC_LO_SCALE, acc[j]);
        }
        if (live) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            reinterpret_cast<float4*>(prow + c0)[qd] = make_float4(acc[4 * qd], acc[4 * qd + 1], acc[4 * qd + 2], acc[4 * qd + 3]);
        }
      }
    } else {
    E4SConv pf = p;                          // the fast form also covers a residual added before the activation (ResNet blocks)
    pf.res = nullptr;
    if (tc_epi_is_fast(pf) && !(p.res && p.res_after_act)) {
      // branch-free epilogue (tc_ptx.cuh): per-row demodulation from global memory (rows of a tile may belong to different
      // regions), the layer-wide bias / slope vectors through a 16-column register block
      const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
      const float nz = er.nz;
      float* optr = p.out + pix * p.out_pitch + n_base;
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 16) {
        float acc[16];
        float4 mul[4], add[4], sl[4];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int n = n_base + c0 + 4 * qd, nl = half * (BN / 2) + c0 + 4 * qd;
          mul[qd] = *reinterpret_cast<const float4*>(sv + nl);                     // channel scale x accumulate un-bias
          if (drow) {
            const float4 dm = ldg4(drow + n);
            mul[qd].x *= dm.x; mul[qd].y *= dm.y; mul[qd].z *= dm.z; mul[qd].w *= dm.w;
          }
          add[qd] = *reinterpret_cast<const float4*>(sv + BN + nl);
          sl[qd] = *reinterpret_cast<const float4*>(sv + 2 * BN + nl);
        }
        if (p.res && live) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const float4 rv = ldg4(p.res + pix * p.res_pitch + n_base + c0 + 4 * qd);
            add[qd].x += rv.x; add[qd].y += rv.y; add[qd].z += rv.z; add[qd].w += rv.w;
          }
        }
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + c0), acc);   // warp-collective
        if (f16) {
          float acc2[16];
          tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + half * (BN / 2) + c0), acc2);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc2[j], 1.f / TC_LO_SCALE, acc[j]);
        }
        if (live) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float4 a;
            a.x = fmaf(acc[4 * qd], mul[qd].x, add[qd].x + nz); a.y = fmaf(acc[4 * qd + 1], mul[qd].y, add[qd].y + nz);
            a.z = fmaf(acc[4 * qd + 2], mul[qd].z, add[qd].z + nz); a.w = fmaf(acc[4 * qd + 3], mul[qd].w, add[qd].w + nz);
            a.x = (a.x < 0.f ? a.x * sl[qd].x : a.x) * gain; a.y = (a.y < 0.f ? a.y * sl[qd].y : a.y) * gain;
            a.z = (a.z < 0.f ? a.z * sl[qd].z : a.z) * gain; a.w = (a.w < 0.f ? a.w * sl[qd].w : a.w) * gain;
            reinterpret_cast<float4*>(optr + c0)[qd] = a;
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 16) {
        float acc[16];
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + c0), acc);   // warp-collective
        if (f16) {
          float acc2[16];
          tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + half * (BN / 2) + c0), acc2);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc2[j], 1.f / TC_LO_SCALE, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] *= corr;
        if (live) tc_epilogue16(p, acc, n_base + c0, er);
      }
    }
    }   // !split-K
    tc_fence_before();
  } else if (warp == TC_PRODUCER_WARPS) {
    // =========================== MMA issuer (whole warp walks the loop, one elected lane issues) ====
    {
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t par = (kc / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, par);
        tc_fence_after();
        const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + TC_A_BYTES;
        const uint32_t b_hi = a_lo + TC_A_BYTES, b_lo = b_hi + B_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t koff = k * 32;  // 16 bf16 = 32 bytes along K inside the swizzle atom
            const uint64_t dah = umma_smem_desc(a_hi + koff), dal = umma_smem_desc(a_lo + koff);
            const uint64_t dbh = umma_smem_desc(b_hi + koff), dbl = umma_smem_desc(b_lo + koff);
            if (f16) {                                             // separate accumulator for the small terms (tc_ptx.cuh)
              umma_bf16(tmem_acc + BN, dal, dbh, IDESC, (kc | k) != 0);
              umma_bf16(tmem_acc + BN, dah, dbl, IDESC, 1);
              umma_bf16(tmem_acc, dah, dbh, IDESC, (kc | k) != 0);
            } else {
              umma_bf16(tmem_acc, dal, dbh, IDESC, (kc | k) != 0);   // small terms first
              umma_bf16(tmem_acc, dah, dbl, IDESC, 1);
              umma_bf16(tmem_acc, dah, dbh, IDESC, 1);
            }
          }
          umma_commit(bar_empty + 8 * s);          // frees this smem stage once the MMAs above have read it
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar_acc);       // accumulator complete -> epilogue
      __syncwarp();
    }
  } else {
    // =========================== B loader (bulk-copy engine) ======================================
    {
      const int phases = up ? 4 : 1;
      const int64_t chunk_bytes = 2 * (int64_t)phases * B_BYTES;             // [hi: phases][lo: phases] per (n_tile, chunk)
      const uint8_t* src = wpk + (int64_t)n_tile * num_kc_all * chunk_bytes + (int64_t)phase_id * B_BYTES;
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t par = (kc / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, par ^ 1);
        const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * TC_A_BYTES;
        if (elect_one()) {
          const int64_t koff = (int64_t)(kc_base + kc) * chunk_bytes;
          mbar_arrive_expect_tx(bar_full + 8 * s, 2 * B_BYTES);
          if (phases == 1) {
            bulk_g2s(dst, src + koff, 2 * B_BYTES, bar_full + 8 * s);   // B_hi | B_lo are contiguous
          } else {
            bulk_g2s(dst, src + koff, B_BYTES, bar_full + 8 * s);
            bulk_g2s(dst + B_BYTES, src + koff + (int64_t)phases * B_BYTES, B_BYTES, bar_full + 8 * s);
          }
        }
        __syncwarp();
      }
    }
  }

  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 2 * BN);
  }
}

// w_f32 [phases][K][cout_pad] -> per (n_tile, k_chunk): [hi tiles of phase 0..P-1 | lo tiles of phase 0..P-1], each tile BN
// rows x 128 bytes; row n holds k = chunk*64 .. +63 (bf16) with the 16-byte chunks XOR-swizzled by (n % 8) == the smem
// image.  Phase-inner, so a kernel that merges the phases of an up-convolution along N (halo: PM = 4, wide: slot pairs)
// fetches a whole K chunk with ONE bulk copy; a single phase is two copies (hi, lo).
__global__ void pack_weights_tc_kernel(const float* __restrict__ w, int phases, int K, int cin, int cout, int cout_pad, int bn,
                                       uint8_t* __restrict__ out, int64_t total, const int fmt, const float wscale) {
  const int num_kc = (K + TC_BK - 1) / TC_BK, nt = cout / bn;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates (phase, n_tile, kc, n_local, kpair) with kpair = 32 bf16 pairs per row
    int kp = (int)(i % 32);
    int64_t t = i / 32;
    int nl = (int)(t % bn);
    t /= bn;
    int kc = (int)(t % num_kc);
    t /= num_kc;
    int ntile = (int)(t % nt);
    int ph = (int)(t / nt);
    // chunk -> k: group-major when cin % 64 == 0 (see conv_tc_kernel), linear otherwise
    const int n = ntile * bn + nl;
    int k;
    if (cin % 64 == 0) {
      const int ntaps = K / cin, g = kc / ntaps, tap = kc - g * ntaps;
      k = tap * cin + g * 64 + kp * 2;
    } else {
      k = kc * TC_BK + kp * 2;
    }
    const float a = k < K ? w[((int64_t)ph * K + k) * cout_pad + n] * wscale : 0.f;      // wscale is a power of two: exact
    const float b = k + 1 < K ? w[((int64_t)ph * K + k + 1) * cout_pad + n] * wscale : 0.f;
    uint32_t h, l;
    tc_split2(fmt == E4S_TC_F16, a, b, h, l);
    const int64_t tile = (((int64_t)ntile * num_kc + kc) * phases) * (2 * (int64_t)bn * 128);     // chunk base: hi[phases] | lo[phases]
    const int chunk = (kp >> 2) ^ (nl & 7);
    const int64_t off = (int64_t)nl * 128 + chunk * 16 + (kp & 3) * 4;
    *reinterpret_cast<uint32_t*>(out + tile + (int64_t)ph * bn * 128 + off) = h;
    *reinterpret_cast<uint32_t*>(out + tile + (int64_t)(phases + ph) * bn * 128 + off) = l;
  }
}

int validate_conv(const E4SConv* p);

static int g_tc_halo = -1;         // -1: read E4S_TC_HALO (default on)
static int g_tc_wide = -1;         // -1: read E4S_TC_WIDE (default on)
bool tc_wide_eligible(const E4SConv* p);
int tc_launch_wide(const E4SConv* p, const void* wpk, cudaStream_t s);
bool tc_halo_eligible(const E4SConv* p);
int tc_launch_halo(const E4SConv* p, const void* wpk, cudaStream_t s, const int4* rjobs, const int* rjob_count, int rjob_host_count);
bool tc_halo_geometry_ok(const E4SConv* p);

static bool tc_shape_ok(int k, int cout) {
  if (k < 8 || k % 8) return false;
  if (cout >= 256) return cout % 256 == 0;
  return cout == 32 || cout == 64 || cout == 128;
}

template <int BN>
static int launch_tc(const E4SConv* p, const void* wpk, int64_t m_total, cudaStream_t s, int ksplit = 1, float* part = nullptr) {
  static bool attr_set_dev[E4S_MAX_DEVICES] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(BN));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(BN));
    if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  dim3 grid((unsigned)ceil_div64(m_total, TC_BM), (unsigned)(p->cout / BN), ksplit > 1 ? (unsigned)ksplit : (up ? 4u : 1u));
  const int gw = up ? p->win : p->wout, gh = up ? p->hin : p->hout;
  const int tile2d = (gw % 8 == 0 && gh % 16 == 0) ? 1 : 0;       // full 16x8 tiles only (tile count is the same)
  if (ksplit > 1) conv_tc_kernel<BN, true><<<grid, TC_THREADS, tc_smem_bytes(BN), s>>>(*p, static_cast<const uint8_t*>(wpk), m_total, tile2d, ksplit, part);
  else conv_tc_kernel<BN, false><<<grid, TC_THREADS, tc_smem_bytes(BN), s>>>(*p, static_cast<const uint8_t*>(wpk), m_total, tile2d, 1, nullptr);
  return check_launch("e4s_conv_tc");
}

}  // namespace e4s

using namespace e4s;

extern "C" int64_t e4s_pack_weights_tc_bytes(int phases, int k, int cout) {
  if (phases < 1 || !tc_shape_ok(k, cout)) return 0;
  return (int64_t)phases * ((k + TC_BK - 1) / TC_BK * TC_BK) * cout * 4;  // hi + lo bf16 per (zero-padded) weight
}

extern "C" int e4s_pack_weights_tc(const float* w_f32, int phases, int k, int cin, int cout, int cout_pad, void* w_packed, void* stream) {
  return e4s_pack_weights_tc_fmt(w_f32, phases, k, cin, cout, cout_pad, E4S_TC_BF16, 1.f, w_packed, stream);
}

extern "C" int e4s_pack_weights_tc_fmt(const float* w_f32, int phases, int k, int cin, int cout, int cout_pad, int fmt, float scale,
                                       void* w_packed, void* stream) {
  E4S_REQUIRE(w_f32 && w_packed, "pack_weights_tc: null pointer");
  E4S_REQUIRE((fmt == E4S_TC_BF16 || fmt == E4S_TC_F16) && scale > 0.f, "pack_weights_tc: bad operand format %d / scale %g", fmt, (double)scale);
  E4S_REQUIRE(phases >= 1 && tc_shape_ok(k, cout) && cout_pad >= cout, "pack_weights_tc: unsupported shape K=%d cout=%d", k, cout);
  E4S_REQUIRE(cin >= 8 && cin % 8 == 0 && k % cin == 0, "pack_weights_tc: K=%d must be taps * cin (cin=%d)", k, cin);
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, "pack_weights_tc: output must be 16-byte aligned");
  const int bn = tc_block_n(cout);
  const int64_t total = (int64_t)phases * (cout / bn) * ((k + TC_BK - 1) / TC_BK) * bn * 32;
  int64_t g = ceil_div64(total, 256);
  if (g > 148 * 32) g = 148 * 32;
  pack_weights_tc_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(w_f32, phases, k, cin, cout, cout_pad, bn, static_cast<uint8_t*>(w_packed), total, fmt, scale);
  return check_launch("pack_weights_tc");
}

namespace e4s {
// out[pix, n] = act( (sum_z part[z][pix][n]) * corr * demod[b, r(pix), n] * ch_scale[n] + noise_w * noise[pix] + ch_shift[n] )   (fast epilogue family)
__global__ void __launch_bounds__(256) conv_splitk_epilogue_kernel(const E4SConv p, const float* __restrict__ part, const int ksplit, const int64_t m_total,
                                                                  const float corr) {
  const int c4 = p.cout >> 2;
  const int64_t total = m_total * c4;
  const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % c4) * 4;
    const int64_t pix = i / c4;
    const int ox = (int)(pix % p.wout);
    const int64_t t = pix / p.wout;
    const int oy = (int)(t % p.hout), b = (int)(t / p.hout);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < ksplit; ++z) {                                   // fixed order: deterministic, batch-invariant
      const float4 v = *reinterpret_cast<const float4*>(part + ((int64_t)z * m_total + pix) * p.cout + n);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    int r = 0;
    if (p.labels) r = p.labels[((int64_t)b * p.lab_h + nearest_src(oy, p.lab_h, p.hout)) * p.lab_w + nearest_src(ox, p.lab_w, p.wout)];
    float4 mul = p.demod ? ldg4(p.demod + ((int64_t)b * p.regions + r) * p.cout + n) : make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.ch_scale) {
      const float4 sc = ldg4(p.ch_scale + n);
      mul.x *= sc.x; mul.y *= sc.y; mul.z *= sc.z; mul.w *= sc.w;
    }
    const float4 add = p.ch_shift ? ldg4(p.ch_shift + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float nz = p.noise ? nw * __ldg(p.noise + (int64_t)b * p.noise_sb + (int64_t)oy * p.wout + ox) : 0.f;
    const float4 sl = make_float4(tc_epi_slope(p, n), tc_epi_slope(p, n + 1), tc_epi_slope(p, n + 2), tc_epi_slope(p, n + 3));
    a.x = fmaf(a.x * corr, mul.x, add.x + nz); a.y = fmaf(a.y * corr, mul.y, add.y + nz);
    a.z = fmaf(a.z * corr, mul.z, add.z + nz); a.w = fmaf(a.w * corr, mul.w, add.w + nz);
    a.x = (a.x < 0.f ? a.x * sl.x : a.x) * gain; a.y = (a.y < 0.f ? a.y * sl.y : a.y) * gain;
    a.z = (a.z < 0.f ? a.z * sl.z : a.z) * gain; a.w = (a.w < 0.f ? a.w * sl.w : a.w) * gain;
    *reinterpret_cast<float4*>(p.out + pix * p.out_pitch + n) = a;
  }
}
}  // namespace e4s

// Split-K form of e4s_conv_tc for layers with too few output tiles to fill the machine (the generator's 4^2 - 16^2 layers: M = batch*16 ...
// batch*256 pixels, K = 4608): ksplit CTAs per output tile each accumulate num_kc / ksplit K chunks, a second kernel adds the partial
// accumulators in index order and applies the epilogue.  Restrictions: mode NORMAL, bf16 split, the piecewise-linear epilogue family
// (tc_epi_is_fast), no fused ToRGB, ksplit divides the number of 64-wide K chunks.  ws: ksplit * batch*hout*wout * cout floats.
extern "C" int64_t e4s_conv_tc_splitk_ws_bytes(const E4SConv* p, int ksplit) {
  if (!p || ksplit < 1) return 0;
  return (int64_t)ksplit * p->batch * p->hout * p->wout * p->cout * 4;
}

extern "C" int e4s_conv_tc_splitk(const E4SConv* p, const void* w_packed, int ksplit, void* ws, void* stream) {
  int rc = validate_conv(p);
  if (rc) return rc;
  E4S_REQUIRE(w_packed && ws && ksplit >= 2 && ksplit <= 64, "conv_tc_splitk: bad args");
  const int K = p->kh * p->kw * p->cin;
  E4S_REQUIRE(tc_shape_ok(K, p->cout), "conv_tc_splitk: needs cin %% 8 == 0 and cout in {32,64,128,256*n} (cin=%d cout=%d)", p->cin, p->cout);
  E4S_REQUIRE(p->mode == E4S_CONV_NORMAL && p->tc_fmt == E4S_TC_BF16 && !p->in_square && !p->rgb && tc_epi_is_fast(*p) && p->out,
              "conv_tc_splitk: unsupported mode / operand format / epilogue option");
  const int num_kc = (K + TC_BK - 1) / TC_BK;
  E4S_REQUIRE(num_kc % ksplit == 0, "conv_tc_splitk: ksplit=%d must divide the %d K chunks", ksplit, num_kc);
  E4S_REQUIRE(p->out_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 && (!p->smod || (reinterpret_cast<uintptr_t>(p->smod) & 15) == 0) &&
                  (!p->demod || (reinterpret_cast<uintptr_t>(p->demod) & 15) == 0),
              "conv_tc_splitk: pointers must be 16-byte aligned, out_pitch %% 4 == 0");
  const int64_t m_total = (int64_t)p->batch * p->hout * p->wout;
  cudaStream_t s = as_stream(stream);
  float* part = static_cast<float*>(ws);
  switch (tc_block_n(p->cout)) {
    case 256: rc = launch_tc<256>(p, w_packed, m_total, s, ksplit, part); break;
    case 128: rc = launch_tc<128>(p, w_packed, m_total, s, ksplit, part); break;
    case 64: rc = launch_tc<64>(p, w_packed, m_total, s, ksplit, part); break;
    case 32: rc = launch_tc<32>(p, w_packed, m_total, s, ksplit, part); break;
    default: return fail(E4S_ERR_UNSUPPORTED, "conv_tc_splitk: unsupported cout %d", p->cout);
  }
  if (rc) return rc;
  const float corr = tc_acc_unbias(*p, (num_kc / ksplit) * (TC_BK / 16) * 3);
  int64_t g = ceil_div64(m_total * (p->cout / 4), 256);
  if (g > 148 * 8) g = 148 * 8;
  conv_splitk_epilogue_kernel<<<(unsigned)g, 256, 0, s>>>(*p, part, ksplit, m_total, corr);
  return check_launch("e4s_conv_tc_splitk(epilogue)");
}

extern "C" int e4s_conv_tc(const E4SConv* p, const void* w_packed, void* stream) {
  int rc = validate_conv(p);
  if (rc) return rc;
  E4S_REQUIRE(w_packed, "conv_tc: null packed weights");
  const int K = p->kh * p->kw * p->cin;
  E4S_REQUIRE(tc_shape_ok(K, p->cout), "conv_tc: needs cin %% 8 == 0 and cout in {32,64,128,256*n} (cin=%d cout=%d)", p->cin, p->cout);
  E4S_REQUIRE(!p->in_square, "conv_tc: in_square is only implemented by the fp32 engine");
  E4S_REQUIRE(p->tc_fmt == E4S_TC_BF16 || p->tc_fmt == E4S_TC_F16, "conv_tc: unknown operand format %d", p->tc_fmt);
  E4S_REQUIRE(p->out_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0, "conv_tc: out must be 16-byte aligned with pitch %% 4 == 0");
  if (p->rgb)
    E4S_REQUIRE(tc_halo_eligible(p) && p->mode != E4S_CONV_UP2_POLYPHASE && p->cout <= 128 && tc_epi_is_fast(*p) && p->regions == 1 &&
                    (reinterpret_cast<uintptr_t>(p->rgb_w) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->rgb_smod) & 15) == 0,
                "conv_tc: the fused ToRGB tail needs an un-masked same-resolution 3x3 layer of the halo kernel (cout <= 128: one n-tile per pixel, piecewise-linear activation)");
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, "conv_tc: packed weights must be 16-byte aligned");
  if (p->smod) E4S_REQUIRE((reinterpret_cast<uintptr_t>(p->smod) & 15) == 0, "conv_tc: smod must be 16-byte aligned");
  if (p->res) E4S_REQUIRE(p->res_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->res) & 15) == 0, "conv_tc: res must be 16-byte aligned with pitch %% 4 == 0");
  if (p->demod) E4S_REQUIRE((reinterpret_cast<uintptr_t>(p->demod) & 15) == 0, "conv_tc: demod must be 16-byte aligned");
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  const int64_t m_total = up ? (int64_t)p->batch * p->hin * p->win : (int64_t)p->batch * p->hout * p->wout;
  cudaStream_t s = as_stream(stream);
  if (g_tc_halo < 0) {
    const char* e = getenv("E4S_TC_HALO");
    g_tc_halo = (e && e[0] == '0') ? 0 : 1;
  }
  const bool f16_up = up && p->tc_fmt == E4S_TC_F16;          // the halo kernel separates the fp16 accumulators on same-resolution layers only
  if ((g_tc_halo || p->rgb) && !f16_up && tc_halo_eligible(p)) return tc_launch_halo(p, w_packed, s, nullptr, nullptr, 0);
  if (g_tc_wide < 0) {
    const char* e = getenv("E4S_TC_WIDE");
    g_tc_wide = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_tc_wide && p->tc_fmt == E4S_TC_BF16 && tc_wide_eligible(p)) return tc_launch_wide(p, w_packed, s);   // the wide kernel is bf16 only
  switch (tc_block_n(p->cout)) {
    case 256: return launch_tc<256>(p, w_packed, m_total, s);
    case 128: return launch_tc<128>(p, w_packed, m_total, s);
    case 64: return launch_tc<64>(p, w_packed, m_total, s);
    case 32: return launch_tc<32>(p, w_packed, m_total, s);
    default: return fail(E4S_ERR_UNSUPPORTED, "conv_tc: unsupported cout %d", p->cout);
  }
}

// Masked (regional) layer through the halo kernel: one job per (16x8 tile, region present) from e4s_region_tile_jobs.
// `job_count` is the device counter the builder filled, `job_count_host` its value (the caller already read it back once
// per forward to choose between this path and the per-row gather kernel).
extern "C" int e4s_conv_tc_regions(const E4SConv* p, const void* w_packed, const int32_t* jobs, const int32_t* job_count, int job_count_host,
                                   void* stream) {
  int rc = validate_conv(p);
  if (rc) return rc;
  E4S_REQUIRE(w_packed && jobs && job_count && job_count_host > 0, "conv_tc_regions: null job list");
  E4S_REQUIRE(p->labels && p->smod, "conv_tc_regions: needs labels and a style table");
  const int K = p->kh * p->kw * p->cin;
  E4S_REQUIRE(tc_shape_ok(K, p->cout) && tc_halo_geometry_ok(p), "conv_tc_regions: geometry not supported by the halo kernel");
  E4S_REQUIRE(!p->in_square && !p->pixw && !p->rgb, "conv_tc_regions: in_square / pixw / fused ToRGB are not supported here");
  E4S_REQUIRE(p->tc_fmt == E4S_TC_BF16 || (p->tc_fmt == E4S_TC_F16 && p->mode != E4S_CONV_UP2_POLYPHASE),
              "conv_tc_regions: operand format %d is not supported for this layer", p->tc_fmt);
  E4S_REQUIRE(p->out_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0, "conv_tc_regions: out must be 16-byte aligned");
  return tc_launch_halo(p, w_packed, as_stream(stream), reinterpret_cast<const int4*>(jobs), job_count, job_count_host);
}

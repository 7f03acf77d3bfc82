// Regional up-convolution at its algorithmic cost: conv_transpose2d(stride 2) as a 2x2 "cell" GEMM on tcgen05 + the 4x4 FIR as a
// separate HBM-bound finishing pass (reference models/stylegan2/model.py:287-300, upfirdn2d_kernel.cu:71-134).
//
// The poly-phase form (conv_tc_wide.cu / conv_tc_halo.cu, UP) folds the blur into four 3x3 phase filters: 36 Cin*Cout MACs per input
// pixel against 9 for the transposed convolution itself.  It was chosen because the mask multiplies AFTER the blur (model.py:395-398):
// every z value inside an output pixel's 4x4 blur window has to be computed with THAT pixel's region style.  Here the same requirement
// is met by materialising z once per (position, region that reads it):
//
//   z[2i+ky, 2j+kx] += x[i,j] * W[ky,kx]      (2H+1 x 2W+1 grid)   is split into cells (cy,cx), 0 <= cy <= H, 0 <= cx <= W, holding the
//   four values z[2cy+py, 2cx+px]; cell (cy,cx) reads the 2x2 input window x[cy+dy, cx+dx], dy,dx in {0,-1}:
//        tap (0,0)  -> all four phases  (W[py][px])          tap (0,-1)  -> phases px=0 (W[py][2])
//        tap (-1,0) -> phases py=0      (W[2][px])           tap (-1,-1) -> phase (0,0) (W[2][2])            = 9 weight blocks, no zeros.
//   A GEMM row is a (cell, region) pair: the regions of the <= 25 output pixels whose blur windows touch the cell (e4s_upz_build_rows;
//   on real face masks 1.1-1.6 rows per cell).  The row's A operand is modulated with its region's style, its four phase results
//   go to Z[row][phase][cout] in fp32, and the finishing pass computes
//        out[q] = act( demod[r(q)] * sum_{u,v} fir'[u,v] * Z[row(cell(q-1+(u,v)), r(q))][phase] + noise + bias )
//   looking the row up through the per-cell (region bit mask, first row) table.  Exact for every one-hot mask; for masks with many
//   regions per cell (per-pixel noise) the row count exceeds the caller's limit and the poly-phase kernel runs instead (device-side
//   launch predicate, no host read-back).
//
// GEMM kernel = the wide kernel's pipeline (8 A-producer warps -> one MMA-issuing thread -> 512 TMEM columns = 4 phase slots x 128
// channels; weights through a cp.async.bulk ring), with rows taken from the list and each tap's MMA covering only the slots it feeds.
#include "tc_ptx.cuh"

namespace e4s {

constexpr int UZ_THREADS = 10 * 32;
constexpr int UZ_A_STAGE = 2 * TC_A_BYTES;            // A hi | lo: 32 KB
constexpr int UZ_SLOT = 128 * 128;                    // 128 weight rows of one 64-wide K chunk (hi or lo): 16 KB
constexpr int UZ_B_STAGE = 2 * UZ_SLOT;               // two slots of hi OR lo weights: 32 KB
constexpr int UZ_B_STAGES = 4;
constexpr int UZ_SMEM = 2 * UZ_A_STAGE + UZ_B_STAGES * UZ_B_STAGE + TC_BM * 16 + 256 + 1024;

// TMEM slot s holds phase (py,px): slot order chosen so that every tap feeds a contiguous slot range
//   slot 0 = (1,0), slot 1 = (0,0), slot 2 = (0,1), slot 3 = (1,1);   tap (0,-1) -> slots 0-1, tap (-1,0) -> slots 1-2, tap (-1,-1) -> slot 1
// The packed weight image (e4s_pack_convt_weights_f32 + e4s_pack_weights_tc with 9 "phases") lists the 9 blocks in the order the
// sub-steps consume them: W[1][0] W[0][0] | W[0][1] W[1][1] | W[1][2] W[0][2] | W[2][0] W[2][1] | W[2][2].
__device__ __constant__ int c_uz_order[9] = {3, 0, 1, 4, 5, 2, 6, 7, 8};

// SW = output channels per CTA = width of one phase slot in TMEM (128; 64 / 32 for the un-masked 512^2 / 1024^2 layers whose cout is that
// small).  rowlist == nullptr: DIRECT mode for un-masked layers -- row i is cell i of the [batch, hin+1, win+1] cell grid, region 0, and
// max_rows is the exact row count (no list, no counter).
template <int SW>
__global__ void __launch_bounds__(UZ_THREADS, 1) conv_tc_upz_kernel(const E4SConv p, const uint8_t* __restrict__ wpk, const int bn_packed,
                                                                    const int2* __restrict__ rowlist, const int* __restrict__ count,
                                                                    const int max_rows, float* __restrict__ z, const int dbg) {
  constexpr int SLOT_B = SW * 128;                    // SW weight rows of one 64-wide K chunk (hi or lo)
  const int nrows = rowlist ? __ldg(count) : max_rows;
  if (nrows > max_rows) return;                                                                    // the list overflowed: the caller's fallback runs
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;
  const int row0 = (int)blockIdx.x * TC_BM;
  if (row0 >= nrows) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int B_OFF = 2 * UZ_A_STAGE;
  int4* rows = reinterpret_cast<int4*>(smem + B_OFF + UZ_B_STAGES * UZ_B_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF + UZ_B_STAGES * UZ_B_STAGE + TC_BM * 16);
  const uint32_t bar_afull = smem_u32(bars);          // 2
  const uint32_t bar_aempty = bar_afull + 16;         // 2
  const uint32_t bar_bfull = bar_aempty + 16;         // 4
  const uint32_t bar_bempty = bar_bfull + 32;         // 4
  const uint32_t bar_acc = bar_bempty + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ngroups = p.cin / 64;
  const int nsteps = 4 * ngroups;                     // A chunks: (64-channel group, tap), group outer

  if (tid < TC_BM) {
    int4 rw = make_int4(-1, 0, 0, 0);
    if (row0 + tid < nrows) {
      if (rowlist) {
        const int2 e = __ldg(rowlist + row0 + tid);
        rw = make_int4(e.x >> 8, e.y >> 16, e.y & 0xffff, e.x & 0xff);    // b, cy, cx, region
      } else {
        const int cwd = p.win + 1, cpp = (p.hin + 1) * cwd, i = row0 + tid;
        const int b = i / cpp, rem = i - b * cpp;
        rw = make_int4(b, rem / cwd, rem % cwd, 0);
      }
    }
    rows[tid] = rw;
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, TC_PRODUCER_WARPS);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < UZ_B_STAGES; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == TC_PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), 4 * SW);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== A producers =====================================================
    const int cg = tid & 7;           // 8-channel group inside the 64-wide K chunk
    const int r0 = tid >> 3;          // rows r0, r0+32, r0+64, r0+96
    int rb[4], ry[4], rx[4], rr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int4 rw = rows[r0 + 32 * i];
      rb[i] = rw.x;
      ry[i] = rw.y;
      rx[i] = rw.z;
      rr[i] = rw.w;
    }
    float4 sreg[4][2];                // modulation of this thread's 8 channels per row: changes every 4 chunks
    // The gather is latency-bound (one chunk = 8 x 16 B per thread, ~1 us from L2): two chunks are kept in flight in registers
    // (buffers 0 / 1 = even / odd steps), so the loads of step+1 travel while step is converted and the MMAs of step-1 run.
    float4 v[2][4][2];
    uint32_t okm[2] = {0u, 0u};
    auto prefetch = [&](int step, const int bufi) {
      const int g = step >> 2, tap = step & 3;
      const int ci = g * 64 + cg * 8;
      const int dy = -(tap >> 1), dx = -(tap & 1);
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int iy = ry[i] + dy, ix = rx[i] + dx;
        const bool ok = rb[i] >= 0 && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win && !(dbg & 4);
        if (ok) {
          const float4* src = reinterpret_cast<const float4*>(p.x + (((int64_t)rb[i] * p.hin + iy) * p.win + ix) * p.x_pitch + ci);
          v[bufi][i][0] = __ldg(src);
          v[bufi][i][1] = __ldg(src + 1);
          m |= 1u << i;
        }
      }
      okm[bufi] = m;
    };
    auto produce = [&](int step, const int bufi) {
      const int s = step & 1;
      mbar_wait(bar_aempty + 8 * s, ((step >> 1) & 1) ^ 1);
      uint8_t* a_hi = smem + s * UZ_A_STAGE;
      uint8_t* a_lo = a_hi + TC_A_BYTES;
      if (p.smod && (step & 3) == 0) {
        const int ci = (step >> 2) * 64 + cg * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (rb[i] >= 0) {
            const float4* sp = reinterpret_cast<const float4*>(p.smod + ((int64_t)rb[i] * p.regions + rr[i]) * p.cin + ci);
            sreg[i][0] = __ldg(sp);
            sreg[i][1] = __ldg(sp + 1);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = r0 + 32 * i;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if ((okm[bufi] >> i) & 1u) {
          f[0] = v[bufi][i][0].x; f[1] = v[bufi][i][0].y; f[2] = v[bufi][i][0].z; f[3] = v[bufi][i][0].w;
          f[4] = v[bufi][i][1].x; f[5] = v[bufi][i][1].y; f[6] = v[bufi][i][1].z; f[7] = v[bufi][i][1].w;
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
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull + 8 * s);
    };

    prefetch(0, 0);
    prefetch(1, 1);
    for (int step = 0; step < nsteps; step += 2) {     // nsteps = 4 * groups: even
      produce(step, 0);
      if (step + 2 < nsteps) prefetch(step + 2, 0);
      produce(step + 1, 1);
      if (step + 3 < nsteps) prefetch(step + 3, 1);
    }

    // =========================== epilogue: Z[row][phase][cout] = acc ================================
    // Each thread owns one accumulator row (TMEM lane); storing it directly would make every warp-wide st.global touch 32 rows (32
    // partial sectors, 2 KB apart).  The warp transposes its 32 rows x 128 columns through shared memory instead (the operand stages
    // are free once bar_acc has completed) and writes each row's 512 bytes with ONE fully coalesced instruction.
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int hh = warp >> 2;               // slot pair (2hh, 2hh + 1)
    const int n_base = (int)blockIdx.y * SW;
    constexpr int F4 = SW / 4;                                                          // float4 per staged row
    constexpr int RPI = 32 / F4;                                                        // rows written per warp instruction (1, 2 or 4)
    float* stage = reinterpret_cast<float*>(smem + warp * (32 * SW * 4));               // 32 rows x SW floats, float4 index ^= row % 8
#pragma unroll 1
    for (int sl = 2 * hh; sl < 2 * hh + 2; ++sl) {
      const int zslot = sl == 0 ? 2 : (sl == 1 ? 0 : (sl == 2 ? 1 : 3));                  // (py,px) -> py*2 + px
      const int taps = sl == 1 ? 4 : (sl == 3 ? 1 : 2);
      const float corr = tc_acc_unbias(p, ngroups * 4 * taps * 3);                        // accumulate steps into this slot
#pragma unroll 1
      for (int c0 = 0; c0 < SW; c0 += 32) {
        if (dbg & 8) break;
        uint32_t acc[32];
        tmem_ld32_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * SW + c0), acc);   // warp-collective
        tmem_wait_ld32(acc);
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          const int f = (c0 >> 2) + qd;
          *reinterpret_cast<uint4*>(stage + lane * SW + ((f ^ (lane & 7)) << 2)) = make_uint4(acc[4 * qd], acc[4 * qd + 1], acc[4 * qd + 2], acc[4 * qd + 3]);
        }
      }
      __syncwarp();
#pragma unroll 4
      for (int r0_ = 0; r0_ < 32; r0_ += RPI) {
        const int r = r0_ + lane / F4, f = lane % F4;                                    // RPI rows per instruction, F4 lanes each
        const int4 rw = rows[q * 32 + r];
        if (rw.x < 0) continue;                                                           // rows beyond the list are at the end of the tile
        float4 a = *reinterpret_cast<const float4*>(stage + r * SW + ((f ^ (r & 7)) << 2));
        a.x *= corr; a.y *= corr; a.z *= corr; a.w *= corr;      // demodulation is applied after the FIR (finishing pass): one region per output pixel
        if ((dbg & 1) && a.x != 123456.789f) continue;
        *reinterpret_cast<float4*>(z + (((int64_t)row0 + q * 32 + r) * 4 + zslot) * p.cout + n_base + 4 * f) = a;
      }
      __syncwarp();
    }
    tc_fence_before();
  } else if (warp == TC_PRODUCER_WARPS) {
    // =========================== MMA issuer ========================================================
    constexpr uint32_t D_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);    // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * SW) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = two slots
    const uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);         // N = one slot
    int bcount = 0;
    for (int step = 0; step < nsteps; ++step) {
      const int s = step & 1;
      const int tap = step & 3;
      mbar_wait(bar_afull + 8 * s, (step >> 1) & 1);
      tc_fence_after();
      const uint32_t a_h = (((smem_base + s * UZ_A_STAGE) >> 4) & 0x3FFFu) | (1u << 16), a_l = a_h + (TC_A_BYTES >> 4);
      const int nsub = tap == 0 ? 2 : 1;
      for (int sub = 0; sub < nsub; ++sub) {
        const uint32_t idesc = tap == 3 ? idesc128 : idesc256;
        const uint32_t col = tap == 0 ? (uint32_t)(sub * 2 * SW) : (tap == 1 ? 0u : (uint32_t)SW);
        const uint32_t tacc = tmem_acc + col;
        // The weights of a sub-step arrive as two ring stages, hi then lo (32 KB each: the ring holds four stages in flight, which
        // hides the ~1 us L2 -> shared-memory latency of a bulk copy behind the MMAs of the stages before it; with two 64 KB stages
        // the copy of sub-step k+2 could only start when the MMAs of sub-step k had finished and the tensor pipe idled a third of
        // the time).  hi stage: A_lo*B_hi + A_hi*B_hi for the four K steps; lo stage: A_hi*B_lo.
#pragma unroll 1
        for (int half = 0; half < 2; ++half, ++bcount) {
          const int bs = bcount & (UZ_B_STAGES - 1);
          mbar_wait(bar_bfull + 8 * bs, (bcount / UZ_B_STAGES) & 1);
          tc_fence_after();
          const uint32_t b_d = (((smem_base + B_OFF + bs * UZ_B_STAGE) >> 4) & 0x3FFFu) | (1u << 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dbg & 2) break;
              if (half == 0) {
                const uint32_t first = (uint32_t)((step | k) != 0);                  // step 0 (group 0, tap (0,0)) writes all four slots first
                asm volatile(
                    "{\n\t"
                    ".reg .pred p, t;\n\t"
                    ".reg .b64 dah, dal, dbh;\n\t"
                    "setp.ne.b32 p, %5, 0;\n\t"
                    "setp.eq.b32 t, 0, 0;\n\t"
                    "mov.b64 dah, {%1, %4};\n\t"
                    "mov.b64 dal, {%2, %4};\n\t"
                    "mov.b64 dbh, {%3, %4};\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %6, p;\n\t"      // small term first
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %6, t;\n\t"
                    "}" ::"r"(tacc),
                    "r"(a_h + 2 * k), "r"(a_l + 2 * k), "r"(b_d + 2 * k), "r"(D_HI), "r"(first), "r"(idesc)
                    : "memory");
              } else {
                asm volatile(
                    "{\n\t"
                    ".reg .pred t;\n\t"
                    ".reg .b64 dah, dbl;\n\t"
                    "setp.eq.b32 t, 0, 0;\n\t"
                    "mov.b64 dah, {%1, %3};\n\t"
                    "mov.b64 dbl, {%2, %3};\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %4, t;\n\t"
                    "}" ::"r"(tacc),
                    "r"(a_h + 2 * k), "r"(b_d + 2 * k), "r"(D_HI), "r"(idesc)
                    : "memory");
              }
            }
            umma_commit(bar_bempty + 8 * bs);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(bar_aempty + 8 * s);
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar_acc);
    __syncwarp();
  } else {
    // =========================== weight loader =====================================================
    // packed chunk (e4s_pack_weights_tc, 9 phases, K = cin): per (n_tile, group): [hi tiles of the 9 blocks | lo tiles], bn_packed rows each
    const int64_t t_bytes = (int64_t)bn_packed * 128;
    const int cb = (int)blockIdx.y;
    const int nt = cb * SW / bn_packed;
    int bcount = 0;
    for (int step = 0; step < nsteps; ++step) {
      const int g = step >> 2, tap = step & 3;
      const int nsub = tap == 0 ? 2 : 1;
      for (int sub = 0; sub < nsub; ++sub) {
        const uint8_t* base = wpk + ((int64_t)nt * ngroups + g) * (18 * t_bytes) + (int64_t)(cb * SW % bn_packed) * 128;
        const int qs = tap == 0 ? sub : tap + 1;           // sub-step 0..4 of the group
        for (int half = 0; half < 2; ++half, ++bcount) {   // hi blocks, then lo blocks (9 block tiles further on)
          const int bs = bcount & (UZ_B_STAGES - 1);
          mbar_wait(bar_bempty + 8 * bs, ((bcount / UZ_B_STAGES) & 1) ^ 1);
          const uint32_t dst = smem_base + B_OFF + bs * UZ_B_STAGE;
          if (elect_one()) {
            if (qs < 4) {                                    // blocks 2qs, 2qs+1
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, 2 * SLOT_B);
              bulk_g2s(dst, base + (9 * half + 2 * qs) * t_bytes, SLOT_B, bar_bfull + 8 * bs);
              bulk_g2s(dst + SLOT_B, base + (9 * half + 2 * qs + 1) * t_bytes, SLOT_B, bar_bfull + 8 * bs);
            } else {                                         // block 8
              mbar_arrive_expect_tx(bar_bfull + 8 * bs, SLOT_B);
              bulk_g2s(dst, base + (9 * half + 8) * t_bytes, SLOT_B, bar_bfull + 8 * bs);
            }
          }
          __syncwarp();
        }
      }
    }
  }

  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 4 * SW);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// (cell, region) row list.  One thread per cell: bit r of the mask = some output pixel whose 4x4 blur window reads one of the cell's
// z values lies in region r.  Rows of a block are contiguous; blocks take their base with one atomicAdd (the order of the blocks in
// the list varies from run to run, the value of every row -- and so the layer's output -- does not).
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upz_build_rows_kernel(const uint8_t* __restrict__ labels, int lab_h, int lab_w, int batch, int hin, int win,
                                                             int2* __restrict__ cells, int2* __restrict__ rowlist, int* __restrict__ count,
                                                             int max_rows) {
  __shared__ int warp_sums[8];
  __shared__ int block_base;
  const int ch = hin + 1, cw = win + 1;
  const int64_t total = (int64_t)batch * ch * cw;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  uint32_t mask = 0;
  int b = 0, cy = 0, cx = 0;
  if (i < total) {
    cx = (int)(i % cw);
    const int64_t t = i / cw;
    cy = (int)(t % ch);
    b = (int)(t / ch);
    const int hout = 2 * hin, wout = 2 * win;
    const int y0 = max(2 * cy - 2, 0), y1 = min(2 * cy + 2, hout - 1), x0 = max(2 * cx - 2, 0), x1 = min(2 * cx + 2, wout - 1);
    for (int oy = y0; oy <= y1; ++oy) {
      const uint8_t* lrow = labels + ((int64_t)b * lab_h + nearest_src(oy, lab_h, hout)) * lab_w;
      for (int ox = x0; ox <= x1; ++ox) mask |= 1u << (lrow[nearest_src(ox, lab_w, wout)] & 31);
    }
  }
  const int n = __popc(mask);
  // block-wide exclusive scan of n
  int incl = n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) {
      const int s = warp_sums[w];
      warp_sums[w] = run;
      run += s;
    }
    block_base = run > 0 ? atomicAdd(count, run) : 0;
  }
  __syncthreads();
  if (i >= total) return;
  const int base = block_base + warp_sums[warp] + incl - n;
  cells[i] = make_int2((int)mask, base);
  int j = 0;
  uint32_t m = mask;
  while (m) {
    const int r = __ffs(m) - 1;
    m &= m - 1;
    if (base + j < max_rows) rowlist[base + j] = make_int2((b << 8) | r, (cy << 16) | cx);
    ++j;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Finishing pass: 4x4 FIR over Z (rows looked up per output pixel's own region), demodulation, noise + bias + activation -> out (NHWC).
// One warp per input pixel (a,b) = the 2x2 output pixels (2a+dy, 2b+dx) and a block of 64 channels (lanes = float2 channel pairs;
// blockIdx.y = channel block).  Those four outputs read the 5x5 z window (2a-1 .. 2a+3) x (2b-1 .. 2b+3) = phases of the 3x3 cells
// around (a,b): when they lie in one region (everywhere except on region boundaries) the window is loaded once, 25 loads for 4
// outputs, all issued before the first use; otherwise each output pixel gathers its own 16 values from the rows of its own region.
// z values outside the (2H+1) x (2W+1) grid are either in non-existent cells (index -1: masked) or phase-1 values of the last cell
// row / column, which the GEMM leaves at exactly 0.  A block walks an 8 x 8 patch of input pixels (one warp per row): the labels of
// a warp's 2 x 16 output pixels are fetched once, and the row lookup of pixel j+1 is in flight while pixel j's window is loaded.
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ldg2(const float* ptr) { return __ldg(reinterpret_cast<const float2*>(ptr)); }

__global__ void __launch_bounds__(256, 2) upz_blur_kernel(const E4SConv p, const float* __restrict__ z, const int2* __restrict__ cells,
                                                          const int* __restrict__ count, const int max_rows, const float* __restrict__ fir) {
  if (__ldg(count) > max_rows) return;
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int px_n = (p.win + 7) >> 3, py_n = (p.hin + 7) >> 3;
  int t = (int)blockIdx.x;
  const int pxi = t % px_n;
  t /= px_n;
  const int pyi = t % py_n;
  const int b = t / py_n;
  const int a = pyi * 8 + warp;
  if (a >= p.hin) return;
  const int c = (int)blockIdx.y * 64 + lane * 2;                        // this lane's channel pair
  const int cw = p.win + 1, ch = p.hin + 1;
  float kf[16];                                                        // upfirdn2d correlates with the flipped kernel
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = __ldg(fir + 15 - i);
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
  const float slope = p.act == E4S_ACT_LRELU ? p.act_slope : 1.f;
  const float2 add = p.ch_shift ? ldg2(p.ch_shift + c) : make_float2(0.f, 0.f);
  const float* dbase = p.demod ? p.demod + (int64_t)b * p.regions * p.cout + c : nullptr;
  const int bx0 = pxi * 8;
  // labels of the warp's 2 x 16 output pixels: lane = dy * 16 + column
  int mylab = 0;
  {
    const int oy = 2 * a + (lane >> 4), ox = 2 * bx0 + (lane & 15);
    if (ox < p.wout) mylab = p.labels[((int64_t)b * p.lab_h + nearest_src(oy, p.lab_h, p.hout)) * p.lab_w + nearest_src(ox, p.lab_w, p.wout)];
  }
  auto row_lookup = [&](const int j) {                                 // lanes 0..8: row of cell (a-1+i, bx-1+k) for the region of pixel j's (0,0) output
    int row = -1;
    const int r = __shfl_sync(0xffffffffu, mylab, 2 * j);
    if (lane < 9 && bx0 + j < p.win) {
      const int cy = a - 1 + lane / 3, cx = bx0 + j - 1 + lane % 3;
      if (cy >= 0 && cx >= 0) {
        const int2 ci = __ldg(cells + ((int64_t)b * ch + cy) * cw + cx);
        row = ci.y + __popc((uint32_t)ci.x & ((1u << r) - 1u));
      }
    }
    return row;
  };
  int myrow_next = row_lookup(0);
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int bx = bx0 + j;
    if (bx >= p.win) break;
    const int myrow = myrow_next;
    const int r00 = __shfl_sync(0xffffffffu, mylab, 2 * j), r01 = __shfl_sync(0xffffffffu, mylab, 2 * j + 1);
    const int r10 = __shfl_sync(0xffffffffu, mylab, 16 + 2 * j), r11 = __shfl_sync(0xffffffffu, mylab, 17 + 2 * j);
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.noise) {
      const float* np = p.noise + (int64_t)b * p.noise_sb + (int64_t)(2 * a) * p.wout + 2 * bx;
      nz[0] = nw * __ldg(np); nz[1] = nw * __ldg(np + 1); nz[2] = nw * __ldg(np + p.wout); nz[3] = nw * __ldg(np + p.wout + 1);
    }
    float* obase = p.out + (((int64_t)b * p.hout + 2 * a) * p.wout + 2 * bx) * p.out_pitch + c;
    if (r00 == r01 && r00 == r10 && r00 == r11) {
      // ---- one region: shared 5x5 window; every load is issued unconditionally (missing cells read row 0 and are masked afterwards)
      const float* zb[9];
      bool zv_ok[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int row = __shfl_sync(0xffffffffu, myrow, i);
        zv_ok[i] = row >= 0;
        zb[i] = z + (int64_t)(row >= 0 ? row : 0) * 4 * p.cout + c;
      }
      float2 zw[25];
#pragma unroll
      for (int wy = 0; wy < 5; ++wy) {
#pragma unroll
        for (int wx = 0; wx < 5; ++wx) {
          const int ci = ((wy + 1) >> 1) * 3 + ((wx + 1) >> 1);        // window index -> cell index 0,1,1,2,2 ; phase 1,0,1,0,1
          zw[wy * 5 + wx] = ldg2(zb[ci] + (((wy + 1) & 1) * 2 + ((wx + 1) & 1)) * p.cout);
        }
      }
      const float2 dm = dbase ? ldg2(dbase + (int64_t)r00 * p.cout) : make_float2(1.f, 1.f);
      if (j + 1 < 8) myrow_next = row_lookup(j + 1);                    // in flight while this pixel's window arrives
      float2 acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
      for (int wy = 0; wy < 5; ++wy) {
#pragma unroll
        for (int wx = 0; wx < 5; ++wx) {
          const int ci = ((wy + 1) >> 1) * 3 + ((wx + 1) >> 1);
          float2 zv = zw[wy * 5 + wx];
          if (!zv_ok[ci]) zv = make_float2(0.f, 0.f);
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int u = wy - dy, v = wx - dx;
              if (u >= 0 && u < 4 && v >= 0 && v < 4) {
                const float k = kf[u * 4 + v];
                acc[dy * 2 + dx].x = fmaf(k, zv.x, acc[dy * 2 + dx].x);
                acc[dy * 2 + dx].y = fmaf(k, zv.y, acc[dy * 2 + dx].y);
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float2 v2;
        v2.x = fmaf(acc[o].x, dm.x, add.x + nz[o]); v2.y = fmaf(acc[o].y, dm.y, add.y + nz[o]);
        v2.x = (v2.x < 0.f ? v2.x * slope : v2.x) * gain; v2.y = (v2.y < 0.f ? v2.y * slope : v2.y) * gain;
        *reinterpret_cast<float2*>(obase + ((int64_t)(o >> 1) * p.wout + (o & 1)) * p.out_pitch) = v2;
      }
    } else {
      // ---- region boundary: every output pixel gathers the rows of its own region -------------
      if (j + 1 < 8) myrow_next = row_lookup(j + 1);
#pragma unroll 1
      for (int o = 0; o < 4; ++o) {
        const int dy = o >> 1, dx = o & 1;
        const int r = o == 0 ? r00 : (o == 1 ? r01 : (o == 2 ? r10 : r11));
        const float nzo = o == 0 ? nz[0] : (o == 1 ? nz[1] : (o == 2 ? nz[2] : nz[3]));
        const int u = (lane >> 2) & 3, v = lane & 3;
        const int zy = 2 * a + dy - 1 + u, zx = 2 * bx + dx - 1 + v;
        int64_t zoff = 0;
        float kc = 0.f;
        if (zy >= 0 && zx >= 0) {                                      // zy <= 2H+1, zx <= 2W+1: cells exist, out-of-grid phases hold 0
          const int2 ci = __ldg(cells + ((int64_t)b * ch + (zy >> 1)) * cw + (zx >> 1));
          const int row = ci.y + __popc((uint32_t)ci.x & ((1u << r) - 1u));
          zoff = ((int64_t)row * 4 + (zy & 1) * 2 + (zx & 1)) * p.cout;
          kc = __ldg(fir + 15 - (u * 4 + v));
        }
        float2 zw[16];
        float kk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int64_t off = __shfl_sync(0xffffffffu, zoff, k);
          kk[k] = __shfl_sync(0xffffffffu, kc, k);
          zw[k] = ldg2(z + off + c);                                   // off 0 (row 0) with coefficient 0 where the window leaves the grid
        }
        const float2 dm = dbase ? ldg2(dbase + (int64_t)r * p.cout) : make_float2(1.f, 1.f);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 zv = kk[k] != 0.f ? zw[k] : make_float2(0.f, 0.f);
          acc.x = fmaf(kk[k], zv.x, acc.x); acc.y = fmaf(kk[k], zv.y, acc.y);
        }
        float2 v2;
        v2.x = fmaf(acc.x, dm.x, add.x + nzo); v2.y = fmaf(acc.y, dm.y, add.y + nzo);
        v2.x = (v2.x < 0.f ? v2.x * slope : v2.x) * gain; v2.y = (v2.y < 0.f ? v2.y * slope : v2.y) * gain;
        *reinterpret_cast<float2*>(obase + ((int64_t)dy * p.wout + dx) * p.out_pitch) = v2;
      }
    }
  }
}

// Finishing pass of the DIRECT (un-masked) mode: row = cell index, one style per sample -- no labels, no lookups.  A thread owns one input
// pixel (a, bx) = 2x2 output pixels and four channels; LPP = min(cout, 128) / 4 threads share a pixel, consecutive thread groups take
// consecutive pixels of a row, so the 25 window loads and the 4 stores of a warp are contiguous runs.
__global__ void __launch_bounds__(256) upz_blur_direct_kernel(const E4SConv p, const float* __restrict__ z, const float* __restrict__ fir,
                                                              const int lpp, const int64_t npix) {
  const int ppb = 256 / lpp;
  const int64_t pixi = (int64_t)blockIdx.x * ppb + threadIdx.x / lpp;
  if (pixi >= npix) return;
  const int cq = (threadIdx.x % lpp) * 4;
  const int bx = (int)(pixi % p.win);
  const int64_t t = pixi / p.win;
  const int a = (int)(t % p.hin), b = (int)(t / p.hin);
  const int cw = p.win + 1, ch = p.hin + 1;
  float kf[16];                                                        // upfirdn2d correlates with the flipped kernel
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = __ldg(fir + 15 - i);
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const float gain = p.act == E4S_ACT_LRELU ? p.act_gain : 1.f;
  const float slope = p.act == E4S_ACT_LRELU ? p.act_slope : 1.f;
  float nz[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.noise) {
    const float* np = p.noise + (int64_t)b * p.noise_sb + (int64_t)(2 * a) * p.wout + 2 * bx;
    nz[0] = nw * __ldg(np); nz[1] = nw * __ldg(np + 1); nz[2] = nw * __ldg(np + p.wout); nz[3] = nw * __ldg(np + p.wout + 1);
  }
  float* obase = p.out + (((int64_t)b * p.hout + 2 * a) * p.wout + 2 * bx) * p.out_pitch;
  const float* zb[9];
  bool ok[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int cy = a - 1 + i / 3, cx = bx - 1 + i % 3;
    ok[i] = cy >= 0 && cx >= 0;                                        // cy <= H, cx <= W always exist; their out-of-grid phases hold 0
    zb[i] = z + (((int64_t)b * ch + (ok[i] ? cy : 0)) * cw + (ok[i] ? cx : 0)) * 4 * p.cout;
  }
  for (int c = cq; c < p.cout; c += 4 * lpp) {
    float4 zw[25];
#pragma unroll
    for (int wy = 0; wy < 5; ++wy) {
#pragma unroll
      for (int wx = 0; wx < 5; ++wx) {
        const int ci = ((wy + 1) >> 1) * 3 + ((wx + 1) >> 1);          // window index -> cell index 0,1,1,2,2 ; phase 1,0,1,0,1
        zw[wy * 5 + wx] = ldg4(zb[ci] + (((wy + 1) & 1) * 2 + ((wx + 1) & 1)) * p.cout + c);
      }
    }
    float4 acc[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int wy = 0; wy < 5; ++wy) {
#pragma unroll
      for (int wx = 0; wx < 5; ++wx) {
        const int ci = ((wy + 1) >> 1) * 3 + ((wx + 1) >> 1);
        float4 zv = zw[wy * 5 + wx];
        if (!ok[ci]) zv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int u = wy - dy, v = wx - dx;
            if (u >= 0 && u < 4 && v >= 0 && v < 4) {
              const float k = kf[u * 4 + v];
              float4& ac = acc[dy * 2 + dx];
              ac.x = fmaf(k, zv.x, ac.x); ac.y = fmaf(k, zv.y, ac.y); ac.z = fmaf(k, zv.z, ac.z); ac.w = fmaf(k, zv.w, ac.w);
            }
          }
        }
      }
    }
    const float4 add = p.ch_shift ? ldg4(p.ch_shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 dm = p.demod ? ldg4(p.demod + (int64_t)b * p.regions * p.cout + c) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float4 v4;
      v4.x = fmaf(acc[o].x, dm.x, add.x + nz[o]); v4.y = fmaf(acc[o].y, dm.y, add.y + nz[o]);
      v4.z = fmaf(acc[o].z, dm.z, add.z + nz[o]); v4.w = fmaf(acc[o].w, dm.w, add.w + nz[o]);
      v4.x = (v4.x < 0.f ? v4.x * slope : v4.x) * gain; v4.y = (v4.y < 0.f ? v4.y * slope : v4.y) * gain;
      v4.z = (v4.z < 0.f ? v4.z * slope : v4.z) * gain; v4.w = (v4.w < 0.f ? v4.w * slope : v4.w) * gain;
      *reinterpret_cast<float4*>(obase + ((int64_t)(o >> 1) * p.wout + (o & 1)) * p.out_pitch + c) = v4;
    }
  }
}

// w [cout][cin][3][3] -> out [9][cin][cout_pad]: block i = scale * W[ky][kx] with ky*3 + kx = c_uz_order[i], as a [cin x cout] matrix
__global__ void __launch_bounds__(256) pack_convt_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int cout_pad,
                                                                 float scale, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    int64_t t = i / cout_pad;
    const int ci = (int)(t % cin);
    const int blk = (int)(t / cin);
    out[i] = co < cout ? scale * w[((int64_t)co * cin + ci) * 9 + c_uz_order[blk]] : 0.f;
  }
}

}  // namespace e4s

using namespace e4s;

static int g_uz_dbg = 0;
/* debug aid (profiling experiments on conv_tc_upz_kernel): bit0 skip the Z stores, bit1 skip the MMAs, bit2 skip the A gather loads, bit3 skip the TMEM reads */
extern "C" int e4s_debug_upz_flags(int flags) {
  g_uz_dbg = flags;
  return E4S_OK;
}

extern "C" int e4s_pack_convt_weights_f32(const float* w, float* out, int cout, int cin, int cout_pad, float scale, void* stream) {
  E4S_REQUIRE(w && out && cout > 0 && cin > 0 && cout_pad >= cout, "pack_convt_weights: bad args");
  const int64_t total = (int64_t)9 * cin * cout_pad;
  int64_t g = ceil_div64(total, 256);
  if (g > 148 * 16) g = 148 * 16;
  pack_convt_weights_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(w, out, cout, cin, cout_pad, scale, total);
  return check_launch("pack_convt_weights");
}

extern "C" int e4s_upz_build_rows(const uint8_t* labels, int batch, int lab_h, int lab_w, int hin, int win, int32_t* cells, int32_t* rows,
                                  int32_t* count, int max_rows, void* stream) {
  E4S_REQUIRE(labels && cells && rows && count && batch > 0 && batch < (1 << 23) && lab_h > 0 && lab_w > 0 && hin > 0 && win > 0 && hin < 32768 &&
                  win < 32768 && max_rows > 0,
              "upz_build_rows: bad args");
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(cells) & 7) == 0 && (reinterpret_cast<uintptr_t>(rows) & 7) == 0, "upz_build_rows: buffers must be 8-byte aligned");
  const int64_t total = (int64_t)batch * (hin + 1) * (win + 1);
  E4S_REQUIRE(total < 0x7fffffff, "upz_build_rows: too many cells");
  upz_build_rows_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(labels, lab_h, lab_w, batch, hin, win,
                                                                                      reinterpret_cast<int2*>(cells), reinterpret_cast<int2*>(rows),
                                                                                      count, max_rows);
  return check_launch("upz_build_rows");
}

template <int SW>
static int launch_upz_gemm(const E4SConv* p, const void* w_packed9, const int32_t* rows, const int32_t* count, int max_rows, float* z, cudaStream_t s) {
  static bool attr_set_dev[E4S_MAX_DEVICES] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_upz_kernel<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, UZ_SMEM);
    if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "conv_tc_upz: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ceil_div(max_rows, TC_BM);
  conv_tc_upz_kernel<SW><<<dim3((unsigned)tiles, (unsigned)(p->cout / SW)), UZ_THREADS, UZ_SMEM, s>>>(
      *p, static_cast<const uint8_t*>(w_packed9), tc_block_n(p->cout), reinterpret_cast<const int2*>(rows), count, max_rows, z, g_uz_dbg);
  return check_launch("e4s_conv_tc_upz(gemm)");
}

extern "C" int e4s_conv_tc_upz(const E4SConv* p, const void* w_packed9, const float* fir, const int32_t* cells, const int32_t* rows,
                               const int32_t* count, int max_rows, float* z, void* stream) {
  E4S_REQUIRE(p && w_packed9 && fir && z && max_rows > 0, "conv_tc_upz: null argument");
  E4S_REQUIRE(p->mode == E4S_CONV_UP2_POLYPHASE && p->kh == 3 && p->kw == 3 && p->hout == 2 * p->hin && p->wout == 2 * p->win,
              "conv_tc_upz: needs the 3x3 up-convolution geometry");
  const bool direct = p->labels == nullptr;              // un-masked layer: one row per cell in cell order, no list
  if (direct) {
    E4S_REQUIRE(!cells && !rows && !count && p->regions == 1 && (int64_t)max_rows == (int64_t)p->batch * (p->hin + 1) * (p->win + 1),
                "conv_tc_upz: the direct (un-masked) form takes no row list, one style per sample and max_rows = batch*(hin+1)*(win+1)");
    E4S_REQUIRE(p->cout == 32 || p->cout == 64 || p->cout % 128 == 0, "conv_tc_upz: the direct form needs cout in {32, 64, 128 n} (cout=%d)", p->cout);
  } else {
    E4S_REQUIRE(cells && rows && count && p->smod && p->regions >= 1 && p->regions <= 32, "conv_tc_upz: needs the row list, a style table and <= 32 regions");
    E4S_REQUIRE(p->cout % 128 == 0, "conv_tc_upz: the regional form needs cout %% 128 == 0 (cout=%d)", p->cout);
  }
  E4S_REQUIRE(p->x && p->out && p->cin % 64 == 0 && p->batch > 0, "conv_tc_upz: needs x, out and cin %% 64 == 0 (cin=%d)", p->cin);
  E4S_REQUIRE(p->tc_fmt == E4S_TC_BF16 && !p->in_mean && !p->in_shift && !p->in_square && !p->rgb && !p->res && !p->pixw && !p->accumulate &&
                  !p->ch_scale && (!p->noise || p->noise_sc == 0) && (p->act == E4S_ACT_NONE || p->act == E4S_ACT_LRELU),
              "conv_tc_upz: unsupported operand format / prologue / epilogue option");
  E4S_REQUIRE(p->x_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && p->out_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (!p->smod || (reinterpret_cast<uintptr_t>(p->smod) & 15) == 0) &&
                  (!p->demod || (reinterpret_cast<uintptr_t>(p->demod) & 15) == 0) && (!p->ch_shift || (reinterpret_cast<uintptr_t>(p->ch_shift) & 15) == 0) &&
                  (reinterpret_cast<uintptr_t>(w_packed9) & 15) == 0,
              "conv_tc_upz: pointers must be 16-byte aligned, pitches %% 4 == 0");
  cudaStream_t s = as_stream(stream);
  int rc;
  if (p->cout == 32) rc = launch_upz_gemm<32>(p, w_packed9, rows, count, max_rows, z, s);
  else if (p->cout == 64) rc = launch_upz_gemm<64>(p, w_packed9, rows, count, max_rows, z, s);
  else rc = launch_upz_gemm<128>(p, w_packed9, rows, count, max_rows, z, s);
  if (rc) return rc;
  if (direct) {
    const int lpp = (p->cout < 128 ? p->cout : 128) / 4;
    const int64_t npix = (int64_t)p->batch * p->hin * p->win;
    const int64_t blocks = ceil_div64(npix, 256 / lpp);
    E4S_REQUIRE(blocks < 0x7fffffff, "conv_tc_upz: too many pixels");
    upz_blur_direct_kernel<<<(unsigned)blocks, 256, 0, s>>>(*p, z, fir, lpp, npix);
    return check_launch("e4s_conv_tc_upz(blur direct)");
  }
  const int64_t patches = (int64_t)p->batch * ((p->hin + 7) / 8) * ((p->win + 7) / 8);
  E4S_REQUIRE(patches < 0x7fffffff, "conv_tc_upz: too many output patches");
  upz_blur_kernel<<<dim3((unsigned)patches, (unsigned)(p->cout / 64)), 256, 0, s>>>(*p, z, reinterpret_cast<const int2*>(cells), count, max_rows, fir);
  return check_launch("e4s_conv_tc_upz(blur)");
}

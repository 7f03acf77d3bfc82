// fp32 CUDA-core implicit-GEMM convolution with the gather/modulate prologue and the fused
// epilogue described in include/e4s_b200.h (E4SConv).  This is the exact-fp32 engine: it
// serves BiSeNet (bit-exact argmax needs fp32-class logits), tiny GEMMs (style -> s, d tables,
// LocalMLP) and any layer shape the tcgen05 engine (conv_tc.cu) does not take.
//
// Tiling: 128 output pixels x BN output channels per CTA, K step 16, 256 threads, 8 x (BN/16)
// accumulators per thread, register-prefetched double-buffered smem tiles.
#include "common.cuh"

namespace e4s {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int APAD = 4;

struct Row {
  int b, oy, ox, r;
};

// no `switch` / long if-chain on `act`: nvcc lowers those to a per-element jump table (LDC + BRX), very slow here
__device__ __forceinline__ float apply_act(float v, int act, float slope, float gain, const float* prelu, int n) {
  if (act <= E4S_ACT_RELU) {   // NONE / LRELU / RELU: (v < 0 ? v*sl : v) * g (+ lower clamp at 0 for RELU)
    const float sl = act == E4S_ACT_LRELU ? slope : 1.f;
    const float g = act == E4S_ACT_LRELU ? gain : 1.f;
    const float lo = act == E4S_ACT_RELU ? 0.f : -INFINITY;
    return fmaxf((v < 0.f ? v * sl : v) * g, lo);
  }
  if (act == E4S_ACT_PRELU) return v < 0.f ? v * __ldg(prelu + n) : v;
  if (act == E4S_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return rsqrtf(v + slope);   // E4S_ACT_RSQRT_EPS
}

template <int BN>
__device__ __forceinline__ void conv_igemm_f32_body(const E4SConv& p, const int64_t m_total, const int bx, const int by, const int phase) {
  constexpr int NH = BN / 64;  // column halves handled per thread (4 columns each)
  if (p.pred_count != nullptr && ((__ldg(p.pred_count) > p.pred_limit) != (p.pred_run_if_gt != 0))) return;   // device-side launch predicate (e4s_b200.h)
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + APAD];
  __shared__ Row rows[BM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int py = phase >> 1, px = phase & 1;
  const int n0 = by * BN;
  const int K = p.kh * p.kw * p.cin;
  const float* wbase = p.w + (int64_t)phase * K * p.cout_pad;

  // ---- row bookkeeping: which output pixel does gather row (tid % 128) belong to -------------
  const int gm = tid & (BM - 1);
  const int ghalf = tid >> 7;
  int gb = -1, goy = 0, gox = 0, gr = 0, ga = 0, gbb = 0;
  {
    int64_t m = (int64_t)bx * BM + gm;
    if (m < m_total) {
      if (p.mode == E4S_CONV_UP2_POLYPHASE) {
        int hw = p.hin * p.win;
        gb = (int)(m / hw);
        int rem = (int)(m - (int64_t)gb * hw);
        ga = rem / p.win;
        gbb = rem - ga * p.win;
        goy = 2 * ga + py;
        gox = 2 * gbb + px;
      } else {
        int hw = p.hout * p.wout;
        gb = (int)(m / hw);
        int rem = (int)(m - (int64_t)gb * hw);
        goy = rem / p.wout;
        gox = rem - goy * p.wout;
      }
      if (p.labels) {
        int sy = nearest_src(goy, p.lab_h, p.hout), sx = nearest_src(gox, p.lab_w, p.wout);
        gr = p.labels[((int64_t)gb * p.lab_h + sy) * p.lab_w + sx];
      }
    }
    if (tid < BM) rows[gm] = Row{gb, goy, gox, gr};
  }
  const float* smod_row = p.smod ? p.smod + ((int64_t)gb * p.regions + gr) * p.cin : nullptr;
  const float* mean_row = p.in_mean ? p.in_mean + (int64_t)gb * p.cin : nullptr;
  const float* rstd_row = p.in_mean ? p.in_rstd + (int64_t)gb * p.cin : nullptr;
  const int hv = p.hin << p.in_shift, wv = p.win << p.in_shift;  // size of the (upsampled) input view

  float4 ra[2];
  float4 rb[NH];

  auto load_tiles = [&](int kt) {
    // A: 8 consecutive k (same tap, cin % 8 == 0) for gather row gm
    int k0 = kt * BK + ghalf * 8;
    ra[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    ra[1] = ra[0];
    if (gb >= 0 && k0 < K) {
      int tap = k0 / p.cin;
      int ci = k0 - tap * p.cin;
      int iy, ix;
      if (p.mode == E4S_CONV_UP2_POLYPHASE) {
        int u = tap / 3;
        iy = ga - 1 + u;
        ix = gbb - 1 + (tap - u * 3);
      } else {
        int ky = tap / p.kw;
        iy = goy * p.stride - p.pad + ky;
        ix = gox * p.stride - p.pad + (tap - ky * p.kw);
      }
      if (iy >= 0 && iy < hv && ix >= 0 && ix < wv) {
        iy >>= p.in_shift;
        ix >>= p.in_shift;
        const float* src = p.x + (((int64_t)gb * p.hin + iy) * p.win + ix) * p.x_pitch + ci;
        ra[0] = __ldg(reinterpret_cast<const float4*>(src));
        ra[1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        if (mean_row) {
          float4 m0 = __ldg(reinterpret_cast<const float4*>(mean_row + ci));
          float4 m1 = __ldg(reinterpret_cast<const float4*>(mean_row + ci) + 1);
          float4 s0 = __ldg(reinterpret_cast<const float4*>(rstd_row + ci));
          float4 s1 = __ldg(reinterpret_cast<const float4*>(rstd_row + ci) + 1);
          ra[0].x = (ra[0].x - m0.x) * s0.x; ra[0].y = (ra[0].y - m0.y) * s0.y;
          ra[0].z = (ra[0].z - m0.z) * s0.z; ra[0].w = (ra[0].w - m0.w) * s0.w;
          ra[1].x = (ra[1].x - m1.x) * s1.x; ra[1].y = (ra[1].y - m1.y) * s1.y;
          ra[1].z = (ra[1].z - m1.z) * s1.z; ra[1].w = (ra[1].w - m1.w) * s1.w;
        }
        if (p.in_square) {
          ra[0].x *= ra[0].x; ra[0].y *= ra[0].y; ra[0].z *= ra[0].z; ra[0].w *= ra[0].w;
          ra[1].x *= ra[1].x; ra[1].y *= ra[1].y; ra[1].z *= ra[1].z; ra[1].w *= ra[1].w;
        }
        if (smod_row) {
          float4 s0 = __ldg(reinterpret_cast<const float4*>(smod_row + ci));
          float4 s1 = __ldg(reinterpret_cast<const float4*>(smod_row + ci) + 1);
          ra[0].x *= s0.x; ra[0].y *= s0.y; ra[0].z *= s0.z; ra[0].w *= s0.w;
          ra[1].x *= s1.x; ra[1].y *= s1.y; ra[1].z *= s1.z; ra[1].w *= s1.w;
        }
      }
    }
    // B: BK x BN weights, float4 along n
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      int idx = tid + i * 256;
      int kr = idx / (BN / 4);
      int c4 = idx - kr * (BN / 4);
      int k = kt * BK + kr;
      int n = n0 + c4 * 4;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K && n < p.cout_pad) rb[i] = __ldg(reinterpret_cast<const float4*>(wbase + (int64_t)k * p.cout_pad + n));
    }
  };
  auto store_tiles = [&](int buf) {
    int kk = ghalf * 8;
    As[buf][kk + 0][gm] = ra[0].x; As[buf][kk + 1][gm] = ra[0].y;
    As[buf][kk + 2][gm] = ra[0].z; As[buf][kk + 3][gm] = ra[0].w;
    As[buf][kk + 4][gm] = ra[1].x; As[buf][kk + 5][gm] = ra[1].y;
    As[buf][kk + 6][gm] = ra[1].z; As[buf][kk + 7][gm] = ra[1].w;
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      int idx = tid + i * 256;
      int kr = idx / (BN / 4);
      int c4 = idx - kr * (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][kr][c4 * 4]) = rb[i];
    }
  };

  float acc[8][NH * 4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NH * 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[NH * 4];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][h * 64 + tx * 4]);
        bv[h * 4 + 0] = b4.x; bv[h * 4 + 1] = b4.y; bv[h * 4 + 2] = b4.z; bv[h * 4 + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NH * 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue ---------------------------------------------------------------------------
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ml = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    const Row rw = rows[ml];
    if (rw.b < 0) continue;
    const int64_t pix = ((int64_t)rw.b * p.hout + rw.oy) * p.wout + rw.ox;
    const float* drow = p.demod ? p.demod + ((int64_t)rw.b * p.regions + rw.r) * p.cout : nullptr;
    float pw = 1.f;
    if (p.pixw) {
      int sy = nearest_src(rw.oy, p.lab_h, p.hout), sx = nearest_src(rw.ox, p.lab_w, p.wout);
      pw = __ldg(p.pixw + (int64_t)rw.b * p.pixw_sb + (int64_t)sy * p.lab_w + sx);
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int nb = n0 + h * 64 + tx * 4;
      if (nb >= p.cout) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        float t = acc[i][h * 4 + j];
        if (n < p.cout) {
          if (drow) t *= __ldg(drow + n);
          if (p.pixw) t *= pw;
          if (p.ch_scale) t *= __ldg(p.ch_scale + n);
          if (p.noise)
            t += nw * __ldg(p.noise + (int64_t)rw.b * p.noise_sb + (int64_t)n * p.noise_sc + (int64_t)rw.oy * p.wout + rw.ox);
          if (p.ch_shift) t += __ldg(p.ch_shift + n);
          if (p.res && !p.res_after_act) t += __ldg(p.res + pix * p.res_pitch + n);
          t = apply_act(t, p.act, p.act_slope, p.act_gain, p.act_prelu, n);
          if (p.res && p.res_after_act) t += __ldg(p.res + pix * p.res_pitch + n);
        }
        v[j] = t;
      }
      float* o = p.out + pix * p.out_pitch + nb;
      if (nb + 3 < p.cout && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
        float4 val = make_float4(v[0], v[1], v[2], v[3]);
        if (p.accumulate) {
          float4 old = *reinterpret_cast<float4*>(o);
          val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
        }
        *reinterpret_cast<float4*>(o) = val;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < p.cout) o[j] = p.accumulate ? o[j] + v[j] : v[j];
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(256) conv_igemm_f32_kernel(const E4SConv p, const int64_t m_total) {
  conv_igemm_f32_body<BN>(p, m_total, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Batched launch of up to CONV_BATCH_MAX independent (mode NORMAL) problems in ONE kernel: blockIdx.z picks the
// problem.  Used for the per-forward style tables (26 modulation + 17 demodulation GEMMs of a Generator forward were
// 43 tiny launches = 8.5 % of the step, profiles/r1_launches_tc_halo.txt).
constexpr int CONV_BATCH_MAX = 32;
struct ConvBatch {
  E4SConv c[CONV_BATCH_MAX];
};

template <int BN>
__global__ void __launch_bounds__(256) conv_igemm_f32_batched_kernel(const __grid_constant__ ConvBatch batch) {
  const E4SConv& p = batch.c[blockIdx.z];
  const int64_t m_total = (int64_t)p.batch * p.hout * p.wout;
  if ((int64_t)blockIdx.x * BM >= m_total || (int)blockIdx.y * BN >= p.cout) return;
  conv_igemm_f32_body<BN>(p, m_total, blockIdx.x, blockIdx.y, 0);
}

// ---- skinny batched linear: y[r, :] = act(scale * (W^T x[r, :]) + shift) for FEW rows against LARGE weight matrices ----------------
// The 12 LocalMLPs (1280 -> 512 -> 6656, reference models/networks.py:23-49) see one row per face: 195 MB of weights against
// 16 rows.  In the 128-row tiles of conv_igemm_f32 that is 0.6 ms per forward (weight reads at 0.3 TB/s); here every CTA streams a
// 128-column slab of W once with float4 loads, K split over the 8 warps (fixed-order shared-memory reduction: deterministic and
// independent of how many rows ride along), 16 rows per pass.  Which kernel a layer takes depends on its (K, N) only, never on the
// row count, so a sample's result does not depend on the batch it is in.
constexpr int SK_ROWS = 16, SK_COLS = 128, SK_WARPS = 8;

__device__ __forceinline__ float sk_act(float v, int act, float slope, float gain) {
  if (act == E4S_ACT_LRELU) return (v < 0.f ? v * slope : v) * gain;
  if (act == E4S_ACT_RELU) return fmaxf(v, 0.f);
  if (act == E4S_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == E4S_ACT_RSQRT_EPS) return rsqrtf(v + slope);
  return v;
}

__global__ void __launch_bounds__(32 * SK_WARPS) linear_skinny_batched_kernel(const __grid_constant__ ConvBatch batch) {
  const E4SConv& p = batch.c[blockIdx.z];
  const int rows = p.batch, K = p.cin;
  const int r0 = blockIdx.y * SK_ROWS, c0 = blockIdx.x * SK_COLS;
  if (r0 >= rows || c0 >= p.cout_pad) return;
  __shared__ float red[SK_ROWS][SK_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = c0 + 4 * lane;
  const bool col_ok = col < p.cout_pad;                       // cout_pad % 4 == 0: whole float4 groups
  const int kper = (K + SK_WARPS - 1) / SK_WARPS;
  const int k_lo = warp * kper, k_hi = min(K, k_lo + kper);
  float4 acc[SK_ROWS];
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nr = min(SK_ROWS, rows - r0);
  const float* xb = p.x + (int64_t)r0 * p.x_pitch;
  if (col_ok) {
#pragma unroll 4
    for (int k = k_lo; k < k_hi; ++k) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(p.w + (int64_t)k * p.cout_pad + col));
#pragma unroll
      for (int r = 0; r < SK_ROWS; ++r) {
        float xv = r < nr ? __ldg(xb + (int64_t)r * p.x_pitch + k) : 0.f;     // warp-uniform address: one broadcast load
        if (p.in_square) xv *= xv;
        acc[r].x = fmaf(xv, wv.x, acc[r].x); acc[r].y = fmaf(xv, wv.y, acc[r].y);
        acc[r].z = fmaf(xv, wv.z, acc[r].z); acc[r].w = fmaf(xv, wv.w, acc[r].w);
      }
    }
  }
  for (int w = 0; w < SK_WARPS; ++w) {                        // warps add their K slices in index order
    if (warp == w) {
#pragma unroll
      for (int r = 0; r < SK_ROWS; ++r) {
        float4* d = reinterpret_cast<float4*>(&red[r][4 * lane]);
        if (w == 0) *d = acc[r];
        else {
          float4 o = *d;
          o.x += acc[r].x; o.y += acc[r].y; o.z += acc[r].z; o.w += acc[r].w;
          *d = o;
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < nr * SK_COLS; i += 32 * SK_WARPS) {
    const int r = i / SK_COLS, c = i - r * SK_COLS, n = c0 + c;
    if (n >= p.cout) continue;
    float v = red[r][c];
    if (p.ch_scale) v *= __ldg(p.ch_scale + n);
    if (p.ch_shift) v += __ldg(p.ch_shift + n);
    p.out[(int64_t)(r0 + r) * p.out_pitch + n] = sk_act(v, p.act, p.act_slope, p.act_gain);
  }
}

// a "rows" problem (linear layer on [rows, K] vectors) whose weight matrix is large enough that streaming it dominates
static bool skinny_problem(const E4SConv& p) {
  return p.mode == E4S_CONV_NORMAL && p.kh == 1 && p.kw == 1 && p.hin == 1 && p.win == 1 && p.stride == 1 && p.pad == 0 && p.in_shift == 0 &&
         !p.in_mean && !p.smod && !p.demod && !p.labels && !p.pixw && !p.noise && !p.res && !p.accumulate && p.act != E4S_ACT_PRELU &&
         (int64_t)p.cin * p.cout >= 512 * 1024;
}

int validate_conv(const E4SConv* p) {
  E4S_REQUIRE(p != nullptr, "conv: null params");
  E4S_REQUIRE(p->x && p->w && (p->out || p->rgb), "conv: null x/w/out");
  if (p->rgb)
    E4S_REQUIRE(p->rgb_w && p->rgb_smod && (!p->rgb_skip || p->rgb_fir) && p->cout % 4 == 0, "conv: fused ToRGB needs rgb_w, rgb_smod (and rgb_fir with a skip)");
  E4S_REQUIRE(p->batch > 0 && p->hin > 0 && p->win > 0 && p->hout > 0 && p->wout > 0, "conv: bad shape");
  E4S_REQUIRE(p->cin > 0 && p->cin % 8 == 0, "conv: cin=%d must be a multiple of 8", p->cin);
  E4S_REQUIRE(p->cout > 0 && p->cout_pad >= p->cout && p->cout_pad % 4 == 0, "conv: bad cout/cout_pad %d/%d", p->cout, p->cout_pad);
  E4S_REQUIRE(p->x_pitch >= p->cin && p->x_pitch % 4 == 0, "conv: x_pitch=%lld", (long long)p->x_pitch);
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              "conv: x / w must be 16-byte aligned");
  E4S_REQUIRE(p->out_pitch >= p->cout, "conv: out_pitch < cout");
  E4S_REQUIRE(p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0 && p->in_shift >= 0, "conv: bad kernel geometry");
  if (p->mode == E4S_CONV_UP2_POLYPHASE) {
    E4S_REQUIRE(p->kh == 3 && p->kw == 3 && p->hout == 2 * p->hin && p->wout == 2 * p->win && p->in_shift == 0,
                "conv: polyphase mode needs 3x3 taps and out = 2*in");
  } else {
    E4S_REQUIRE(p->mode == E4S_CONV_NORMAL, "conv: unknown mode %d", p->mode);
    int hv = p->hin << p->in_shift, wv = p->win << p->in_shift;
    E4S_REQUIRE(p->hout == (hv + 2 * p->pad - p->kh) / p->stride + 1 && p->wout == (wv + 2 * p->pad - p->kw) / p->stride + 1,
                "conv: hout/wout inconsistent with geometry");
  }
  E4S_REQUIRE(!p->in_mean || p->in_rstd, "conv: in_mean without in_rstd");
  if (p->smod || p->demod) E4S_REQUIRE(p->regions > 0, "conv: regions must be > 0 with smod/demod");
  if (p->labels || p->pixw) E4S_REQUIRE(p->lab_h > 0 && p->lab_w > 0, "conv: labels/pixw need lab_h/lab_w");
  E4S_REQUIRE(!p->noise || p->noise_w, "conv: noise without noise_w");
  E4S_REQUIRE(p->act != E4S_ACT_PRELU || p->act_prelu, "conv: PReLU without slopes");
  E4S_REQUIRE(p->act >= 0 && p->act <= E4S_ACT_RSQRT_EPS, "conv: unknown act %d", p->act);
  return E4S_OK;
}

}  // namespace e4s

extern "C" int e4s_conv_f32(const E4SConv* p, void* stream) {
  using namespace e4s;
  int rc = validate_conv(p);
  if (rc) return rc;
  E4S_REQUIRE(!p->rgb, "conv_f32: the fused ToRGB tail is implemented by the tensor-core halo kernel only");
  const bool up = p->mode == E4S_CONV_UP2_POLYPHASE;
  const int64_t m_total = up ? (int64_t)p->batch * p->hin * p->win : (int64_t)p->batch * p->hout * p->wout;
  const int64_t mt = ceil_div64(m_total, BM);
  E4S_REQUIRE(mt <= 0x7fffffff, "conv: too many tiles");
  if (p->cout > 64) {
    dim3 grid((unsigned)mt, (unsigned)ceil_div(p->cout, 128), up ? 4 : 1);
    conv_igemm_f32_kernel<128><<<grid, 256, 0, as_stream(stream)>>>(*p, m_total);
  } else {
    dim3 grid((unsigned)mt, 1, up ? 4 : 1);
    conv_igemm_f32_kernel<64><<<grid, 256, 0, as_stream(stream)>>>(*p, m_total);
  }
  return check_launch("e4s_conv_f32");
}

extern "C" int e4s_conv_f32_batched(const E4SConv* params, int count, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(params && count > 0, "conv_f32_batched: empty batch");
  cudaStream_t s = as_stream(stream);
  for (int base = 0; base < count; base += CONV_BATCH_MAX) {
    const int n = count - base < CONV_BATCH_MAX ? count - base : CONV_BATCH_MAX;
    ConvBatch batch;
    int64_t max_mt = 1;
    int max_cout = 1, max_rows = 1;
    bool skinny = true;
    for (int i = 0; i < n; ++i) {
      int rc = validate_conv(&params[base + i]);
      if (rc) return rc;
      E4S_REQUIRE(params[base + i].mode == E4S_CONV_NORMAL, "conv_f32_batched: only E4S_CONV_NORMAL problems");
      batch.c[i] = params[base + i];
      const int64_t mt = ceil_div64((int64_t)params[base + i].batch * params[base + i].hout * params[base + i].wout, BM);
      if (mt > max_mt) max_mt = mt;
      if (params[base + i].cout > max_cout) max_cout = params[base + i].cout;
      if (params[base + i].batch > max_rows) max_rows = params[base + i].batch;
      skinny = skinny && skinny_problem(params[base + i]);
    }
    E4S_REQUIRE(max_mt <= 0x7fffffff, "conv_f32_batched: too many tiles");
    if (skinny) {                                              // decided by the layers' (K, N) only: batch-invariant results
      dim3 grid((unsigned)ceil_div(max_cout, SK_COLS), (unsigned)ceil_div(max_rows, SK_ROWS), (unsigned)n);
      linear_skinny_batched_kernel<<<grid, 32 * SK_WARPS, 0, s>>>(batch);
      int rc = check_launch("e4s_conv_f32_batched(skinny)");
      if (rc) return rc;
      continue;
    }
    dim3 grid((unsigned)max_mt, (unsigned)ceil_div(max_cout, 128), (unsigned)n);
    conv_igemm_f32_batched_kernel<128><<<grid, 256, 0, s>>>(batch);
    int rc = check_launch("e4s_conv_f32_batched");
    if (rc) return rc;
  }
  return E4S_OK;
}

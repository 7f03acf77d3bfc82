// Reductions and gating of the encoder / BiSeNet: per-(sample,channel) statistics (InstanceNorm,
// global average pooling), tiny per-sample FC layers (SE, attention), the fused
// normalise * gate + shortcut tail, masked region means and 3x3/2 max pooling.  All NHWC fp32,
// all deterministic (fixed reduction trees, fp64 partial sums) so results do not depend on batch size.
#include "common.cuh"

namespace e4s {

constexpr int STATS_MAX_SPLIT = 32;

__host__ __device__ inline int stats_split(int batch, int hw, int c) {
  int groups = batch * ((c + 31) / 32);
  int s = (592 + groups - 1) / groups;           // ~4 blocks per SM (12 measured slower: 13.2 vs 12.8 ms for the encoder stage)
  int cap = hw / 64;
  if (s > cap) s = cap;
  if (s > STATS_MAX_SPLIT) s = STATS_MAX_SPLIT;
  return s < 1 ? 1 : s;
}

// partial[b][s][c] = (sum, sumsq) over the s-th slice of pixels
__global__ void __launch_bounds__(256) chan_stats_partial_kernel(const float* __restrict__ x, int64_t pitch, int hw, int c,
                                                                 int split, double2* __restrict__ partial) {
  __shared__ double2 red[8][32];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, s = blockIdx.y, b = blockIdx.z;
  const int per = (hw + split - 1) / split;
  const int p0 = s * per, p1 = min(hw, p0 + per);
  const int ch = c0 + cx;
  double sum = 0.0, sq = 0.0;
  if (ch < c) {
    const float* xp = x + (int64_t)b * hw * pitch + ch;
    for (int p = p0 + py; p < p1; p += 8) {
      float v = __ldg(xp + (int64_t)p * pitch);
      sum += (double)v;
      sq += (double)v * (double)v;
    }
  }
  red[py][cx] = make_double2(sum, sq);
  __syncthreads();
  if (py == 0 && ch < c) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      sum += red[i][cx].x;
      sq += red[i][cx].y;
    }
    partial[((int64_t)b * split + s) * c + ch] = make_double2(sum, sq);
  }
}

__global__ void chan_stats_final_kernel(const double2* __restrict__ partial, int batch, int hw, int c, int split, float eps,
                                        float* __restrict__ mean, float* __restrict__ rstd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * c) return;
  int b = i / c, ch = i - b * c;
  double sum = 0.0, sq = 0.0;
  for (int s = 0; s < split; ++s) {
    double2 v = partial[((int64_t)b * split + s) * c + ch];
    sum += v.x;
    sq += v.y;
  }
  double m = sum / hw;
  double var = sq / hw - m * m;
  if (var < 0.0) var = 0.0;
  if (mean) mean[i] = (float)m;
  if (rstd) rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// The same statistics in ONE launch when the tensor allows 16-byte loads (c % 4 == 0, pitch % 4 == 0): each thread streams four channels
// (float4 per pixel, 32 pixel rows per block pass), the block writes its slice's partial sums, and the LAST block of a (sample, 32-channel
// group) to arrive -- a ticket counter in the first 64 KB of the workspace, which must be zero when the call starts and is zero again when
// it ends -- reduces the slices in index order and writes mean / rstd.  The result does not depend on which block happens to be last.
constexpr int STATS_CNT_BYTES = 65536;
__global__ void __launch_bounds__(256) chan_stats_fused_kernel(const float* __restrict__ x, int64_t pitch, int hw, int c, int split, int batch, float eps,
                                                               double2* __restrict__ partial, int* __restrict__ counters, float* __restrict__ mean,
                                                               float* __restrict__ rstd) {
  __shared__ double red[32][8][8];                         // [pixel row][4-channel group][sum x4 | sumsq x4]
  __shared__ int s_last;
  const int cq = threadIdx.x & 7, py = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 32, sl = blockIdx.y, b = blockIdx.z;
  const int per = (hw + split - 1) / split;
  const int p0 = sl * per, p1 = min(hw, p0 + per);
  const int ch = c0 + cq * 4;
  double sm[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
  if (ch < c) {
    const float* xp = x + (int64_t)b * hw * pitch + ch;
#pragma unroll 4
    for (int p = p0 + py; p < p1; p += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xp + (int64_t)p * pitch));
      sm[0] += (double)v.x; sq[0] += (double)v.x * (double)v.x;
      sm[1] += (double)v.y; sq[1] += (double)v.y * (double)v.y;
      sm[2] += (double)v.z; sq[2] += (double)v.z * (double)v.z;
      sm[3] += (double)v.w; sq[3] += (double)v.w * (double)v.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[py][cq][j] = sm[j];
    red[py][cq][4 + j] = sq[j];
  }
  __syncthreads();
  if (threadIdx.x < 32 && c0 + threadIdx.x < c) {
    const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
    double a = 0.0, q = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      a += red[r][g][j];
      q += red[r][g][4 + j];
    }
    partial[((int64_t)b * split + sl) * c + c0 + threadIdx.x] = make_double2(a, q);
  }
  __threadfence();
  __syncthreads();
  int* cnt = counters + (int64_t)b * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) s_last = atomicAdd(cnt, 1) == split - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32 && c0 + threadIdx.x < c) {
    const int chn = c0 + threadIdx.x;
    double a = 0.0, q = 0.0;
    for (int k = 0; k < split; ++k) {
      const double2 v = __ldcg(&partial[((int64_t)b * split + k) * c + chn]);
      a += v.x;
      q += v.y;
    }
    const double m = a / hw;
    double var = q / hw - m * m;
    if (var < 0.0) var = 0.0;
    if (mean) mean[(int64_t)b * c + chn] = (float)m;
    if (rstd) rstd[(int64_t)b * c + chn] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (threadIdx.x == 0) *cnt = 0;
}

// one warp per output feature
__global__ void __launch_bounds__(256) vec_fc_kernel(const float* __restrict__ x, int64_t x_stride,
                                                     const float* __restrict__ w, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, float* __restrict__ y, int cin,
                                                     int cout, int act) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= cout) return;
  const float* xr = x + (int64_t)b * x_stride;
  const float* wr = w + (int64_t)o * cin;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) acc = fmaf(__ldg(wr + i), __ldg(xr + i), acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    if (scale) acc *= __ldg(scale + o);
    if (shift) acc += __ldg(shift + o);
    if (act == E4S_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == E4S_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    y[(int64_t)b * cout + o] = acc;
  }
}

__global__ void residual_combine_kernel(const float* __restrict__ a, int64_t a_pitch, const float* __restrict__ a_mean,
                                        const float* __restrict__ a_rstd, const float* __restrict__ gate, int gate_plus_one,
                                        const float* __restrict__ r, int64_t r_pitch, int r_h, int r_w, int r_sub,
                                        const float* __restrict__ r_mean, const float* __restrict__ r_rstd, int relu,
                                        const float* __restrict__ prelu, float* __restrict__ out, int64_t out_pitch, int h, int w, int c, int64_t total4) {
  const int c4 = c >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % c4);
    const int64_t pix = i / c4;
    const int ch = cq * 4;
    const int hw = h * w;
    const int b = (int)(pix / hw);
    const int rem = (int)(pix - (int64_t)b * hw);
    float4 v = __ldg(reinterpret_cast<const float4*>(a + pix * a_pitch + ch));
    if (a_mean) {
      float4 m = __ldg(reinterpret_cast<const float4*>(a_mean + (int64_t)b * c + ch));
      float4 s = __ldg(reinterpret_cast<const float4*>(a_rstd + (int64_t)b * c + ch));
      v.x = (v.x - m.x) * s.x; v.y = (v.y - m.y) * s.y; v.z = (v.z - m.z) * s.z; v.w = (v.w - m.w) * s.w;
    }
    if (gate) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gate + (int64_t)b * c + ch));
      if (gate_plus_one) {  // feat*g + feat, evaluated like the reference (mul then add)
        v.x = v.x * g.x + v.x; v.y = v.y * g.y + v.y; v.z = v.z * g.z + v.z; v.w = v.w * g.w + v.w;
      } else {
        v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
      }
    }
    if (r) {
      const int y = rem / w, xx = rem - y * w;
      int ry, rx;
      if (r_sub > 1) {
        ry = y * r_sub;
        rx = xx * r_sub;
      } else {
        ry = nearest_src(y, r_h, h);
        rx = nearest_src(xx, r_w, w);
      }
      float4 q = __ldg(reinterpret_cast<const float4*>(r + (((int64_t)b * r_h + ry) * r_w + rx) * r_pitch + ch));
      if (r_mean) {
        float4 m = __ldg(reinterpret_cast<const float4*>(r_mean + (int64_t)b * c + ch));
        float4 s = __ldg(reinterpret_cast<const float4*>(r_rstd + (int64_t)b * c + ch));
        q.x = (q.x - m.x) * s.x; q.y = (q.y - m.y) * s.y; q.z = (q.z - m.z) * s.z; q.w = (q.w - m.w) * s.w;
      }
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (prelu) {
      float4 sl = __ldg(reinterpret_cast<const float4*>(prelu + ch));
      v.x = v.x < 0.f ? v.x * sl.x : v.x; v.y = v.y < 0.f ? v.y * sl.y : v.y;
      v.z = v.z < 0.f ? v.z * sl.z : v.z; v.w = v.w < 0.f ? v.w * sl.w : v.w;
    }
    *reinterpret_cast<float4*>(out + pix * out_pitch + ch) = v;
  }
}

constexpr int MM_MAXK = 16;

constexpr int MM_PY = 16;          // pixel rows per CTA (512 threads): the kernel is a chain of L2 round trips per pixel step

__global__ void __launch_bounds__(32 * MM_PY) masked_mean_kernel(const float* __restrict__ feat, int64_t f_pitch, int h, int w,
                                                                 int c, const float* __restrict__ mask, int k, int mh, int mw,
                                                                 float* __restrict__ codes, int64_t sb, int64_t sk, int c_off,
                                                                 int k_total, int k0) {
  __shared__ double red[MM_PY][32];
  __shared__ int redc[MM_PY];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, b = blockIdx.y;
  const int ch = c0 + cx;
  const int hw = h * w;
  double acc[MM_MAXK];
  int cnt[MM_MAXK];
#pragma unroll
  for (int j = 0; j < MM_MAXK; ++j) {
    acc[j] = 0.0;
    cnt[j] = 0;
  }
  const float* mb = mask + ((int64_t)b * k_total + k0) * mh * mw;     // regions k0 .. k0 + k - 1 of this sample
  // Two pixels per step, all 2k mask taps and both feature values loaded BEFORE any of them is tested: the first
  // version branched on each mask load (k dependent L2 round trips per pixel, 1.4 ms per call at 64^2 x 256 channels).
  // Each thread still adds its pixels in increasing order and adding 0.0 for the regions a pixel is not in leaves the sums
  // unchanged, so the result does not depend on how pixels are grouped into steps.
  for (int p = py; p < hw; p += 2 * MM_PY) {
    const int p1 = p + MM_PY;
    const bool has1 = p1 < hw;
    const int y0 = p / w, x0 = p - y0 * w;
    const int y1 = has1 ? p1 / w : y0, x1 = has1 ? p1 - y1 * w : x0;
    const int sy0 = nearest_src(y0, mh, h), sx0 = nearest_src(x0, mw, w);
    const int sy1 = nearest_src(y1, mh, h), sx1 = nearest_src(x1, mw, w);
    const float v0 = ch < c ? __ldg(feat + ((int64_t)b * hw + p) * f_pitch + ch) : 0.f;
    const float v1 = (ch < c && has1) ? __ldg(feat + ((int64_t)b * hw + p1) * f_pitch + ch) : 0.f;
    float m0[MM_MAXK], m1[MM_MAXK];
#pragma unroll
    for (int j = 0; j < MM_MAXK; ++j) {
      m0[j] = j < k ? __ldg(mb + ((int64_t)j * mh + sy0) * mw + sx0) : 0.f;
      m1[j] = (j < k && has1) ? __ldg(mb + ((int64_t)j * mh + sy1) * mw + sx1) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < MM_MAXK; ++j) {
      const bool in0 = m0[j] != 0.f, in1 = m1[j] != 0.f;
      acc[j] += in0 ? (double)v0 : 0.0;
      acc[j] += in1 ? (double)v1 : 0.0;
      cnt[j] += (in0 ? 1 : 0) + (in1 ? 1 : 0);
    }
  }
#pragma unroll
  for (int j = 0; j < MM_MAXK; ++j) {
    if (j >= k) break;
    __syncthreads();
    red[py][cx] = acc[j];
    if (cx == 0) redc[py] = cnt[j];
    __syncthreads();
    if (py == 0 && ch < c) {
      double s = 0.0;
      int n = 0;
#pragma unroll
      for (int i = 0; i < MM_PY; ++i) {
        s += red[i][cx];
        n += redc[i];
      }
      codes[(int64_t)b * sb + (int64_t)(k0 + j) * sk + c_off + ch] = n > 0 ? (float)(s / n) : 0.f;
    }
  }
}

// ---- per-region masked mean through a region-membership bit map -------------------------------------------------------------
// The encoder pools three feature maps (64^2, 32^2, 16^2) with the same 512^2 mask.  Reading K float planes per feature pixel made
// masked_mean_kernel a chain of L2 round trips (0.58 ms at 64^2 x 256 channels, 1.8 % of the copy bandwidth).  Here the mask is
// reduced ONCE per forward to a u32 membership map (bit j set <=> mask[b,j,y,x] != 0: psp_encoders.py:363 tests `mask != 0`-style
// membership, so soft / overlapping masks keep their exact semantics) and every pooling call reads one word per pixel; pixels are
// split into MMB_SPLIT chunks per (sample, 32-channel slab) with fp64 partials reduced in a fixed order (batch-invariant).
__global__ void __launch_bounds__(256) mask_member_bits_kernel(const float* __restrict__ mask, int k, int64_t hw, uint32_t* __restrict__ bits,
                                                               int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw, p = i - b * hw;
    const float* m = mask + b * k * hw + p;
    uint32_t v = 0;
    for (int j = 0; j < k; ++j) v |= (__ldg(m + (int64_t)j * hw) != 0.f ? 1u : 0u) << j;
    bits[i] = v;
  }
}

constexpr int MMB_SPLIT = 8;       // pixel chunks per (sample, channel slab)
constexpr int MMB_MAXK = 32;

// partial: grid (c/32, MMB_SPLIT, batch), 256 threads = 8 pixel lanes x 32 channels; ws_sum [b][slab][chunk][k][32] doubles, ws_cnt ints
template <int KU>
__global__ void __launch_bounds__(256) masked_mean_bits_partial_kernel(const float* __restrict__ feat, int64_t f_pitch, int h, int w, int c,
                                                                       const uint32_t* __restrict__ bits, int k, int mh, int mw,
                                                                       double* __restrict__ ws_sum, int* __restrict__ ws_cnt) {
  __shared__ double red[8][32];
  __shared__ int redc[8];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int slab = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
  const int ch = slab * 32 + cx;
  const int hw = h * w;
  const int per = (hw + MMB_SPLIT - 1) / MMB_SPLIT;
  const int p_lo = chunk * per, p_hi = min(hw, p_lo + per);
  double acc[KU];
  int cnt[KU];
#pragma unroll
  for (int j = 0; j < KU; ++j) {
    acc[j] = 0.0;
    cnt[j] = 0;
  }
  const uint32_t* bb = bits + (int64_t)b * mh * mw;
  for (int p = p_lo + py; p < p_hi; p += 16) {          // two pixels per step, loads first
    const int p1 = p + 8;
    const bool has1 = p1 < p_hi;
    const int y0 = p / w, x0 = p - y0 * w;
    const int y1 = has1 ? p1 / w : y0, x1 = has1 ? p1 - y1 * w : x0;
    const uint32_t m0 = __ldg(bb + (int64_t)nearest_src(y0, mh, h) * mw + nearest_src(x0, mw, w));
    const uint32_t m1 = has1 ? __ldg(bb + (int64_t)nearest_src(y1, mh, h) * mw + nearest_src(x1, mw, w)) : 0u;
    const float v0 = ch < c ? __ldg(feat + ((int64_t)b * hw + p) * f_pitch + ch) : 0.f;
    const float v1 = (ch < c && has1) ? __ldg(feat + ((int64_t)b * hw + p1) * f_pitch + ch) : 0.f;
#pragma unroll
    for (int j = 0; j < KU; ++j) {
      const bool in0 = (m0 >> j) & 1u, in1 = (m1 >> j) & 1u;
      acc[j] += in0 ? (double)v0 : 0.0;
      acc[j] += in1 ? (double)v1 : 0.0;
      cnt[j] += (in0 ? 1 : 0) + (in1 ? 1 : 0);
    }
  }
  const int64_t base = (((int64_t)b * gridDim.x + slab) * MMB_SPLIT + chunk) * k;
#pragma unroll
  for (int j = 0; j < KU; ++j) {
    if (j >= k) break;
    __syncthreads();
    red[py][cx] = acc[j];
    if (cx == 0) redc[py] = cnt[j];
    __syncthreads();
    if (py == 0) {
      double s = 0.0;
      int n = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s += red[i][cx];
        n += redc[i];
      }
      ws_sum[(base + j) * 32 + cx] = s;
      if (cx == 0) ws_cnt[base + j] = n;
    }
  }
}

// final: one thread per (b, region, channel): chunks summed in index order
__global__ void __launch_bounds__(256) masked_mean_bits_final_kernel(const double* __restrict__ ws_sum, const int* __restrict__ ws_cnt, int c, int k,
                                                                     int slabs, float* __restrict__ codes, int64_t sb, int64_t sk, int c_off,
                                                                     int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int j = (int)((i / c) % k);
    const int b = (int)(i / ((int64_t)c * k));
    const int slab = ch >> 5, cx = ch & 31;
    double s = 0.0;
    int n = 0;
    for (int q = 0; q < MMB_SPLIT; ++q) {
      const int64_t base = (((int64_t)b * slabs + slab) * MMB_SPLIT + q) * k + j;
      s += ws_sum[base * 32 + cx];
      n += ws_cnt[base];
    }
    codes[(int64_t)b * sb + (int64_t)j * sk + c_off + ch] = n > 0 ? (float)(s / n) : 0.f;
  }
}

__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, int h, int w, int c, int ho, int wo, float* __restrict__ y,
                                    int64_t total4) {
  const int c4 = c >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % c4);
    int64_t t = i / c4;
    const int ox = (int)(t % wo);
    t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * h + iy) * w + ix) * c + cq * 4));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    *reinterpret_cast<float4*>(y + i * 4) = m;
  }
}

static inline unsigned grid_for(int64_t n, int block) {
  int64_t g = ceil_div64(n, block);
  const int64_t cap = 148 * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace e4s

using namespace e4s;

extern "C" int64_t e4s_chan_stats_ws_bytes(int batch, int c) {
  return STATS_CNT_BYTES + (int64_t)batch * STATS_MAX_SPLIT * c * (int64_t)sizeof(double2);
}

extern "C" int e4s_chan_stats_f32(const float* x, int64_t x_pitch, int batch, int hw, int c, float eps, float* mean,
                                  float* rstd, void* ws, void* stream) {
  E4S_REQUIRE(x && ws && (mean || rstd), "chan_stats: null pointer");
  E4S_REQUIRE(batch > 0 && hw > 0 && c > 0 && x_pitch >= c, "chan_stats: bad shape");
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "chan_stats: workspace must be 16-byte aligned");
  const int split = stats_split(batch, hw, c);
  dim3 grid(ceil_div(c, 32), split, batch);
  double2* partial = reinterpret_cast<double2*>(static_cast<uint8_t*>(ws) + STATS_CNT_BYTES);
  if (c % 4 == 0 && x_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (int64_t)batch * grid.x * 4 <= STATS_CNT_BYTES) {
    chan_stats_fused_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_pitch, hw, c, split, batch, eps, partial, static_cast<int*>(ws), mean, rstd);
    return check_launch("chan_stats_fused");
  }
  chan_stats_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_pitch, hw, c, split, partial);
  int rc = check_launch("chan_stats_partial");
  if (rc) return rc;
  chan_stats_final_kernel<<<ceil_div(batch * c, 256), 256, 0, as_stream(stream)>>>(partial, batch, hw, c, split, eps, mean, rstd);
  return check_launch("chan_stats_final");
}

extern "C" int e4s_vec_fc_f32(const float* x, int64_t x_stride, const float* w, const float* scale, const float* shift,
                              float* y, int batch, int cin, int cout, int act, void* stream) {
  E4S_REQUIRE(x && w && y && batch > 0 && cin > 0 && cout > 0, "vec_fc: bad args");
  E4S_REQUIRE(act == E4S_ACT_NONE || act == E4S_ACT_RELU || act == E4S_ACT_SIGMOID, "vec_fc: unsupported act %d", act);
  dim3 grid(ceil_div(cout, 8), batch);
  vec_fc_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_stride, w, scale, shift, y, cin, cout, act);
  return check_launch("vec_fc");
}

extern "C" int e4s_residual_combine_f32(const float* a, int64_t a_pitch, const float* a_mean, const float* a_rstd,
                                        const float* gate, int gate_plus_one, const float* r, int64_t r_pitch, int r_h,
                                        int r_w, int r_sub, const float* r_mean, const float* r_rstd, int relu, const float* prelu,
                                        float* out, int64_t out_pitch, int batch, int h, int w, int c, void* stream) {
  E4S_REQUIRE(a && out && batch > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "residual_combine: bad args");
  E4S_REQUIRE(a_pitch % 4 == 0 && out_pitch % 4 == 0 && (!r || r_pitch % 4 == 0), "residual_combine: pitches must be multiples of 4");
  E4S_REQUIRE(!a_mean || a_rstd, "residual_combine: a_mean without a_rstd");
  E4S_REQUIRE(!r || (r_h > 0 && r_w > 0 && r_sub >= 1), "residual_combine: bad residual geometry");
  E4S_REQUIRE(!r || r_sub == 1 || ((h - 1) * r_sub < r_h && (w - 1) * r_sub < r_w), "residual_combine: subsampled residual out of range");
  int64_t total4 = (int64_t)batch * h * w * (c / 4);
  residual_combine_kernel<<<grid_for(total4, 256), 256, 0, as_stream(stream)>>>(a, a_pitch, a_mean, a_rstd, gate, gate_plus_one, r,
                                                                                r_pitch, r_h, r_w, r_sub, r_mean, r_rstd, relu,
                                                                                prelu, out, out_pitch, h, w, c, total4);
  return check_launch("residual_combine");
}

extern "C" int e4s_masked_mean_f32(const float* feat, int64_t f_pitch, int batch, int h, int w, int c, const float* mask, int k,
                                   int mh, int mw, float* codes, int64_t codes_stride_b, int64_t codes_stride_k, int c_off,
                                   void* stream) {
  E4S_REQUIRE(feat && mask && codes && batch > 0 && h > 0 && w > 0 && c > 0, "masked_mean: bad args");
  E4S_REQUIRE(k > 0 && mh > 0 && mw > 0, "masked_mean: bad mask shape");
  dim3 grid(ceil_div(c, 32), batch);
  int rc = E4S_OK;
  for (int k0 = 0; k0 < k && rc == E4S_OK; k0 += MM_MAXK) {       // MM_MAXK regions per launch (the 19-class UI variant takes two)
    const int kn = k - k0 < MM_MAXK ? k - k0 : MM_MAXK;
    masked_mean_kernel<<<grid, 32 * MM_PY, 0, as_stream(stream)>>>(feat, f_pitch, h, w, c, mask, kn, mh, mw, codes, codes_stride_b,
                                                           codes_stride_k, c_off, k, k0);
    rc = check_launch("masked_mean");
  }
  return rc;
}

extern "C" int e4s_mask_member_bits_u32(const float* mask, int batch, int k, int h, int w, uint32_t* bits, void* stream) {
  E4S_REQUIRE(mask && bits && batch > 0 && k > 0 && k <= MMB_MAXK && h > 0 && w > 0, "mask_member_bits: bad args (k must be in 1..32)");
  const int64_t hw = (int64_t)h * w, total = (int64_t)batch * hw;
  mask_member_bits_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(mask, k, hw, bits, total);
  return check_launch("mask_member_bits");
}

extern "C" int64_t e4s_masked_mean_ws_bytes(int batch, int c, int k) {
  const int64_t n = (int64_t)batch * ceil_div(c, 32) * MMB_SPLIT * k;
  return n * 32 * (int64_t)sizeof(double) + n * (int64_t)sizeof(int) + 256;
}

extern "C" int e4s_masked_mean_bits_f32(const float* feat, int64_t f_pitch, int batch, int h, int w, int c, const uint32_t* bits, int k,
                                        int mh, int mw, float* codes, int64_t codes_stride_b, int64_t codes_stride_k, int c_off, void* ws,
                                        void* stream) {
  E4S_REQUIRE(feat && bits && codes && ws && batch > 0 && h > 0 && w > 0 && c > 0, "masked_mean_bits: bad args");
  E4S_REQUIRE(k > 0 && k <= MMB_MAXK && mh > 0 && mw > 0, "masked_mean_bits: k must be in 1..%d", MMB_MAXK);
  E4S_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "masked_mean_bits: workspace must be 16-byte aligned");
  const int slabs = ceil_div(c, 32);
  const int64_t n = (int64_t)batch * slabs * MMB_SPLIT * k;
  double* ws_sum = static_cast<double*>(ws);
  int* ws_cnt = reinterpret_cast<int*>(static_cast<char*>(ws) + ((n * 32 * (int64_t)sizeof(double) + 15) & ~(int64_t)15));
  dim3 grid(slabs, MMB_SPLIT, batch);
  if (k <= 16)
    masked_mean_bits_partial_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(feat, f_pitch, h, w, c, bits, k, mh, mw, ws_sum, ws_cnt);
  else
    masked_mean_bits_partial_kernel<32><<<grid, 256, 0, as_stream(stream)>>>(feat, f_pitch, h, w, c, bits, k, mh, mw, ws_sum, ws_cnt);
  int rc = check_launch("masked_mean_bits_partial");
  if (rc) return rc;
  const int64_t total = (int64_t)batch * k * c;
  masked_mean_bits_final_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ws_sum, ws_cnt, c, k, slabs, codes, codes_stride_b,
                                                                                    codes_stride_k, c_off, total);
  return check_launch("masked_mean_bits_final");
}

extern "C" int e4s_maxpool3x3s2_nhwc_f32(const float* x, int batch, int h, int w, int c, float* y, void* stream) {
  E4S_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "maxpool: bad args");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  int64_t total4 = (int64_t)batch * ho * wo * (c / 4);
  maxpool3x3s2_kernel<<<grid_for(total4, 256), 256, 0, as_stream(stream)>>>(x, h, w, c, ho, wo, y, total4);
  return check_launch("maxpool3x3s2");
}

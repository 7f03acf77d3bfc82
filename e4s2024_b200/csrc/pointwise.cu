// Bandwidth-bound kernels of the hot path: layout adapters, upfirdn2d, bias+leaky-ReLU, mask -> label
// maps, ToRGB (+bias +FIR-upsampled skip), one-hot.  All are coalesced / smem-staged streaming kernels;
// grids are sized from the problem (many waves on 148 SMs).
#include "common.cuh"

namespace e4s {

// ------------------------------------------------------------------------------------------------
// layout adapters (module API is NCHW, the engine is NHWC)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int c, int hw, int c_pad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int cc = c0 + i, pp = p0 + tx;
    tile[i][tx] = (cc < c && pp < hw) ? x[((int64_t)b * c + cc) * hw + pp] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int pp = p0 + i, cc = c0 + tx;
    if (pp < hw && cc < c_pad) y[((int64_t)b * hw + pp) * c_pad + cc] = tile[tx][i];
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int64_t pitch, float* __restrict__ y, int c, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    int pp = p0 + i, cc = c0 + tx;
    tile[i][tx] = (pp < hw && cc < c) ? x[((int64_t)b * hw + pp) * pitch + cc] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int cc = c0 + i, pp = p0 + tx;
    if (cc < c && pp < hw) y[((int64_t)b * c + cc) * hw + pp] = tile[tx][i];
  }
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Tiled variant for down == 1, up in {1,2}, taps <= 4x4 (the two modes the generator uses):
// a 16x64 output tile per CTA, input tile + flipped taps staged in shared memory.
template <int UP>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                             float* __restrict__ out, int in_h, int in_w, int out_h,
                                                             int out_w, int kh, int kw, int pad_x0, int pad_y0) {
  constexpr int TH = 16, TW = 64;
  constexpr int IH = (TH + 3) / UP + 2, IW = (TW + 3) / UP + 2;
  __shared__ float sk[4][4];
  __shared__ float sx[IH][IW + 1];
  const int plane = blockIdx.z;
  const int oy0 = blockIdx.y * TH, ox0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  if (tid < 16) {
    int ky = tid >> 2, kx = tid & 3;
    sk[ky][kx] = (ky < kh && kx < kw) ? k[(kh - 1 - ky) * kw + (kw - 1 - kx)] : 0.f;  // flipped taps
  }
  const int iy0 = floor_div(oy0 - pad_y0, UP), ix0 = floor_div(ox0 - pad_x0, UP);
  const float* xp = x + (int64_t)plane * in_h * in_w;
  for (int i = tid; i < IH * IW; i += 256) {
    int ly = i / IW, lx = i - ly * IW;
    int iy = iy0 + ly, ix = ix0 + lx;
    sx[ly][lx] = (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) ? __ldg(xp + (int64_t)iy * in_w + ix) : 0.f;
  }
  __syncthreads();
  const int tx = tid & 63, tyb = tid >> 6;
  const int ox = ox0 + tx;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int oy = oy0 + tyb + 4 * j;
    if (oy >= out_h || ox >= out_w) continue;
    const int my = oy - pad_y0, mx = ox - pad_x0;  // window origin in the zero-inserted image
    float v = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int uy = my + ky;
      if (UP == 2 && (uy & 1)) continue;
      const int ly = (UP == 2 ? (uy >> 1) : uy) - iy0;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ux = mx + kx;
        if (UP == 2 && (ux & 1)) continue;
        const int lx = (UP == 2 ? (ux >> 1) : ux) - ix0;
        v = fmaf(sx[ly][lx], sk[ky][kx], v);
      }
    }
    out[((int64_t)plane * out_h + oy) * out_w + ox] = v;
  }
}

// Generic variant (any up/down/pad, any tap count): one thread per output, taps through L1.
__global__ void upfirdn2d_generic_kernel(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ out,
                                         int64_t total, int in_h, int in_w, int out_h, int out_w, int kh, int kw, int up_x,
                                         int up_y, int down_x, int down_y, int pad_x0, int pad_y0) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ox = (int)(i % out_w);
    int64_t t = i / out_w;
    int oy = (int)(t % out_h);
    int64_t plane = t / out_h;
    const float* xp = x + plane * in_h * in_w;
    const int my = oy * down_y - pad_y0, mx = ox * down_x - pad_x0;
    float v = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      int uy = my + ky;
      if (uy < 0 || uy % up_y) continue;
      int iy = uy / up_y;
      if (iy >= in_h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        int ux = mx + kx;
        if (ux < 0 || ux % up_x) continue;
        int ix = ux / up_x;
        if (ix >= in_w) continue;
        v = fmaf(__ldg(xp + (int64_t)iy * in_w + ix), __ldg(k + (kh - 1 - ky) * kw + (kw - 1 - kx)), v);
      }
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// bias + leaky ReLU (fused_bias_act case 30)
// ------------------------------------------------------------------------------------------------
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y,
                                int64_t n, int64_t inner, int channels, float slope, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (bias) v += __ldg(bias + (i / inner) % channels);
    y[i] = (v < 0.f ? v * slope : v) * scale;
  }
}

__global__ void bias_act_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ bias, float4* __restrict__ y,
                                     int64_t n4, int64_t inner4, int channels, float slope, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    float b = bias ? __ldg(bias + (i / inner4) % channels) : 0.f;  // inner % 4 == 0: one channel per float4
    v.x += b; v.y += b; v.z += b; v.w += b;
    v.x = (v.x < 0.f ? v.x * slope : v.x) * scale;
    v.y = (v.y < 0.f ? v.y * slope : v.y) * scale;
    v.z = (v.z < 0.f ? v.z * slope : v.z) * scale;
    v.w = (v.w < 0.f ? v.w * slope : v.w) * scale;
    y[i] = v;
  }
}

__global__ void noise_bias_act_nhwc_kernel(float* __restrict__ x, int64_t total, int hw, int w_, int c,
                                           const float* __restrict__ noise, const float* __restrict__ noise_w,
                                           int64_t nsb, int64_t nsc, const float* __restrict__ bias, float slope,
                                           float scale) {
  const float nw = noise ? __ldg(noise_w) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    int64_t pix = i / c;
    int b = (int)(pix / hw);
    int p = (int)(pix - (int64_t)b * hw);
    float v = x[i];
    if (noise) v += nw * __ldg(noise + b * nsb + ch * nsc + p);
    if (bias) v += __ldg(bias + ch);
    x[i] = (v < 0.f ? v * slope : v) * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// mask -> label map, one-hot
// ------------------------------------------------------------------------------------------------
__global__ void mask_labels_kernel(const float* __restrict__ mask, int k, int64_t hw, int64_t total,
                                   uint8_t* __restrict__ labels, int32_t* __restrict__ flags) {
  int bad = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / hw, p = i - b * hw;
    const float* m = mask + b * k * hw + p;
    int ones = 0, nonzero = 0, idx = k;
    float best = 0.f;
    for (int j = 0; j < k; ++j) {
      float v = __ldg(m + (int64_t)j * hw);
      if (v != 0.f) {
        ++nonzero;
        if (v == 1.f) ++ones;
        if (idx == k || v > best) { idx = j; best = v; }
      }
    }
    if (!(nonzero == 1 && ones == 1)) bad = 1;
    labels[i] = (uint8_t)(idx < k ? idx : 0);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(flags, 1);
}

__global__ void labels_to_onehot_kernel(const uint8_t* __restrict__ labels, int k, int64_t hw, int64_t total,
                                        float* __restrict__ onehot) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i % hw;
    int64_t t = i / hw;
    int j = (int)(t % k);
    int64_t b = t / k;
    onehot[i] = (labels[b * hw + p] == j) ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// ToRGB: 1x1 modulated conv (no demod) + bias + FIR-upsampled skip, NHWC in -> NCHW out.
// A CTA owns TORGB_TILE consecutive pixels.  Phase A: LP lanes cooperate on one pixel (coalesced float4 channel
// slices, warp-shuffle reduction) and park the three dot products in shared memory.  Phase B: one thread per pixel
// adds bias and the 2x2 non-zero taps of the zero-insert-upsampled skip (closed form of upfirdn2d(up=2, pad=(2,1)))
// and writes the three NCHW planes fully coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int TORGB_TILE_MAX = 256;

// TAB (large images, cin <= 256, regions <= 12, tiles that lie inside one sample): the CTA first builds the MODULATED weights
// wm[region][3][cin] = wrgb * smod[b, region] of its sample in shared memory, so that a pixel step loads only its features from global
// memory (one float4 per lane and channel step instead of five: features, style row and three weight rows) -- the kernel was
// load-issue bound at 1.5 TB/s on the 128^2 / 256^2 layers.
constexpr int TORGB_TAB_R = 12, TORGB_TAB_C = 256;
template <int LP, int TORGB_TILE, bool TAB = false>
__global__ void __launch_bounds__(256) torgb_kernel(const float* __restrict__ x, int64_t x_pitch, int64_t npix, int h, int w,
                                                    int cin, const float* __restrict__ smod, const float* __restrict__ wrgb,
                                                    const uint8_t* __restrict__ labels, int regions, int lab_h, int lab_w,
                                                    const float* __restrict__ pixw, int64_t pixw_sb,
                                                    const float* __restrict__ bias, const float* __restrict__ skip,
                                                    const float* __restrict__ fir, float* __restrict__ rgb, int accumulate) {
  constexpr int PPW = 32 / LP;  // pixels per warp step
  __shared__ float srgb[3][TORGB_TILE];
  __shared__ __align__(16) float wm[TAB ? TORGB_TAB_R * 3 * TORGB_TAB_C : 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LP, ll = lane % LP;
  const int hw = h * w;
  const int64_t tile0 = (int64_t)blockIdx.x * TORGB_TILE;
  if (TAB) {
    const int b0 = (int)(tile0 / hw);                  // the host guarantees hw % TORGB_TILE == 0: one sample per tile
    const int n = regions * 3 * cin;
    for (int idx = threadIdx.x; idx < n; idx += 256) {
      const int r = idx / (3 * cin), rem = idx - r * 3 * cin;
      const int c = rem / cin, ci = rem - c * cin;
      wm[idx] = __ldg(wrgb + c * cin + ci) * __ldg(smod + ((int64_t)b0 * regions + r) * cin + ci);
    }
    __syncthreads();
  }
  // ---- phase A: dot products ---------------------------------------------------------------------
  // TU pixels per warp step with independent accumulators: every step is a chain of dependent global loads (label ->
  // style row -> features), and one pixel per step made the kernel pure latency (90 us per launch whatever the grid size).
  constexpr int TU = 2;
  constexpr int NCI = 4;               // channel steps issued together: all their loads are in flight before the first FMA
  for (int t0 = warp * PPW + sub; t0 < TORGB_TILE; t0 += 8 * PPW * TU) {       // whole warp iterates together (t differs by sub only)
    float a0[TU], a1[TU], a2[TU];
    const float* xr[TU];
    const float* sr[TU];
    int rg[TU];
    bool ok[TU];
#pragma unroll
    for (int u = 0; u < TU; ++u) {
      const int t = t0 + u * 8 * PPW;
      const int64_t pix = tile0 + t;
      a0[u] = a1[u] = a2[u] = 0.f;
      ok[u] = t < TORGB_TILE && pix < npix;
      xr[u] = x;
      sr[u] = smod;
      rg[u] = 0;
      if (ok[u]) {
        const int b = (int)(pix / hw);
        int r = 0;
        if (labels) {
          const int rem = (int)(pix - (int64_t)b * hw);
          const int y = rem / w, xx = rem - y * w;
          r = labels[((int64_t)b * lab_h + nearest_src(y, lab_h, h)) * lab_w + nearest_src(xx, lab_w, w)];
        }
        xr[u] = x + pix * x_pitch;
        sr[u] = smod + ((int64_t)b * regions + r) * cin;
        rg[u] = r;
      }
    }
    if (TAB) {
      for (int ci0 = ll * 4; ci0 < cin; ci0 += LP * 4 * NCI) {
        float4 v[TU][NCI];
#pragma unroll
        for (int k = 0; k < NCI; ++k) {
          const int ci = ci0 + k * LP * 4;
#pragma unroll
          for (int u = 0; u < TU; ++u)
            v[u][k] = (ci < cin && ok[u]) ? __ldg(reinterpret_cast<const float4*>(xr[u] + ci)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NCI; ++k) {
          const int ci = ci0 + k * LP * 4;
          if (ci >= cin) continue;
#pragma unroll
          for (int u = 0; u < TU; ++u) {
            const float* wr = wm + rg[u] * 3 * cin + ci;
            const float4 q = v[u][k];
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + cin),
                         w2 = *reinterpret_cast<const float4*>(wr + 2 * cin);
            a0[u] += q.x * w0.x + q.y * w0.y + q.z * w0.z + q.w * w0.w;
            a1[u] += q.x * w1.x + q.y * w1.y + q.z * w1.z + q.w * w1.w;
            a2[u] += q.x * w2.x + q.y * w2.y + q.z * w2.z + q.w * w2.w;
          }
        }
      }
    } else
    for (int ci0 = ll * 4; ci0 < cin; ci0 += LP * 4 * NCI) {
      float4 w0[NCI], w1[NCI], w2[NCI], v[TU][NCI], sm[TU][NCI];
#pragma unroll
      for (int k = 0; k < NCI; ++k) {
        const int ci = ci0 + k * LP * 4;
        const bool in = ci < cin;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        w0[k] = in ? __ldg(reinterpret_cast<const float4*>(wrgb + ci)) : z;
        w1[k] = in ? __ldg(reinterpret_cast<const float4*>(wrgb + cin + ci)) : z;
        w2[k] = in ? __ldg(reinterpret_cast<const float4*>(wrgb + 2 * cin + ci)) : z;
#pragma unroll
        for (int u = 0; u < TU; ++u) {
          v[u][k] = (in && ok[u]) ? __ldg(reinterpret_cast<const float4*>(xr[u] + ci)) : z;
          sm[u][k] = (in && ok[u]) ? __ldg(reinterpret_cast<const float4*>(sr[u] + ci)) : z;
        }
      }
#pragma unroll
      for (int k = 0; k < NCI; ++k)          // same order of additions per pixel as one channel step at a time
#pragma unroll
        for (int u = 0; u < TU; ++u) {
          // x * (w * s), the association of the TAB path (whose table holds w * s): which path a launch takes depends on the batch
          // size, and a sample must come out bit-identical alone and inside a batch
          const float4 q = v[u][k], s4 = sm[u][k];
          const float4 m0 = make_float4(w0[k].x * s4.x, w0[k].y * s4.y, w0[k].z * s4.z, w0[k].w * s4.w);
          const float4 m1 = make_float4(w1[k].x * s4.x, w1[k].y * s4.y, w1[k].z * s4.z, w1[k].w * s4.w);
          const float4 m2 = make_float4(w2[k].x * s4.x, w2[k].y * s4.y, w2[k].z * s4.z, w2[k].w * s4.w);
          a0[u] += q.x * m0.x + q.y * m0.y + q.z * m0.z + q.w * m0.w;
          a1[u] += q.x * m1.x + q.y * m1.y + q.z * m1.z + q.w * m1.w;
          a2[u] += q.x * m2.x + q.y * m2.y + q.z * m2.z + q.w * m2.w;
        }
    }
#pragma unroll
    for (int u = 0; u < TU; ++u) {
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) {
        a0[u] += __shfl_xor_sync(0xffffffffu, a0[u], o);
        a1[u] += __shfl_xor_sync(0xffffffffu, a1[u], o);
        a2[u] += __shfl_xor_sync(0xffffffffu, a2[u], o);
      }
      const int t = t0 + u * 8 * PPW;
      if (ll == 0 && t < TORGB_TILE) {
        srgb[0][t] = a0[u];
        srgb[1][t] = a1[u];
        srgb[2][t] = a2[u];
      }
    }
  }
  __syncthreads();
  // ---- phase B: bias + skip + coalesced NCHW store -------------------------------------------------
  const int t = threadIdx.x;
  const int64_t pix = tile0 + t;
  if (t >= TORGB_TILE || pix >= npix) return;
  const int b = (int)(pix / hw);
  const int rem = (int)(pix - (int64_t)b * hw);
  const int y = rem / w, xx = rem - y * w;
  float pw = 1.f;
  if (pixw) pw = __ldg(pixw + (int64_t)b * pixw_sb + (int64_t)nearest_src(y, lab_h, h) * lab_w + nearest_src(xx, lab_w, w));
  // skip taps: zero-inserted row u = y - 2 + ky is non-zero only for even u  ->  ky in {y&1, (y&1)+2}
  const int sh = h >> 1, sw = w >> 1;
  const int ky0 = y & 1, kx0 = xx & 1;
  const int sy0 = (y - 2 + ky0) >> 1, sx0 = (xx - 2 + kx0) >> 1;     // arithmetic shift: -1 for the top/left border
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = srgb[c][t] * pw;
    float* o = rgb + ((int64_t)b * 3 + c) * hw + rem;
    if (accumulate) {
      *o += v;
      continue;
    }
    if (bias) v += __ldg(bias + c);
    if (skip) {
      const float* sp = skip + ((int64_t)b * 3 + c) * sh * sw;
      float u = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int sy = sy0 + a, ky = ky0 + 2 * a;
        if (sy < 0 || sy >= sh) continue;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int sx = sx0 + d, kx = kx0 + 2 * d;
          if (sx < 0 || sx >= sw) continue;
          u = fmaf(__ldg(sp + (int64_t)sy * sw + sx), __ldg(fir + (3 - ky) * 4 + (3 - kx)), u);
        }
      }
      v += u;
    }
    *o = v;
  }
}

static inline unsigned grid_for(int64_t n, int block, int per_thread = 1) {
  int64_t g = ceil_div64(n, (int64_t)block * per_thread);
  const int64_t cap = 148 * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace e4s

using namespace e4s;

extern "C" int e4s_nchw_to_nhwc_f32(const float* x, float* y, int batch, int c, int h, int w, int c_pad, void* stream) {
  E4S_REQUIRE(x && y && batch > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "nchw_to_nhwc: bad args");
  int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c_pad, 32), batch);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, y, c, hw, c_pad);
  return check_launch("nchw_to_nhwc");
}

extern "C" int e4s_nhwc_to_nchw_f32(const float* x, int64_t x_pitch, float* y, int batch, int c, int h, int w, void* stream) {
  E4S_REQUIRE(x && y && batch > 0 && c > 0 && h > 0 && w > 0 && x_pitch >= c, "nhwc_to_nchw: bad args");
  int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), batch);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, x_pitch, y, c, hw);
  return check_launch("nhwc_to_nchw");
}

extern "C" int e4s_upfirdn2d_f32(const float* x, const float* kernel, float* out, int64_t planes, int in_h, int in_w,
                                 int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                                 int pad_y0, int pad_y1, void* stream) {
  E4S_REQUIRE(x && kernel && out, "upfirdn2d: null pointer");
  E4S_REQUIRE(planes > 0 && in_h > 0 && in_w > 0 && kh > 0 && kw > 0, "upfirdn2d: bad shape");
  E4S_REQUIRE(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0, "upfirdn2d: bad up/down");
  // same shape rule as upfirdn2d_kernel.cu:167-168 of the reference
  const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  E4S_REQUIRE(out_h > 0 && out_w > 0, "upfirdn2d: empty output (%d x %d)", out_h, out_w);
  cudaStream_t s = as_stream(stream);
  const bool tiled = down_x == 1 && down_y == 1 && up_x == up_y && (up_x == 1 || up_x == 2) && kh <= 4 && kw <= 4 &&
                     planes <= 65535;
  if (tiled) {
    dim3 grid(ceil_div(out_w, 64), ceil_div(out_h, 16), (unsigned)planes);
    if (up_x == 1)
      upfirdn2d_tile_kernel<1><<<grid, 256, 0, s>>>(x, kernel, out, in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0);
    else
      upfirdn2d_tile_kernel<2><<<grid, 256, 0, s>>>(x, kernel, out, in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0);
  } else {
    int64_t total = planes * out_h * out_w;
    upfirdn2d_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, kernel, out, total, in_h, in_w, out_h, out_w, kh, kw,
                                                                 up_x, up_y, down_x, down_y, pad_x0, pad_y0);
  }
  return check_launch("upfirdn2d");
}

extern "C" int e4s_bias_act_f32(const float* x, const float* bias, float* y, int64_t n, int64_t inner, int channels,
                                float slope, float scale, void* stream) {
  E4S_REQUIRE(x && y && n > 0 && inner > 0 && channels > 0, "bias_act: bad args");
  cudaStream_t s = as_stream(stream);
  const bool vec = (inner % 4 == 0) && (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (vec)
    bias_act_vec4_kernel<<<grid_for(n / 4, 256, 2), 256, 0, s>>>(reinterpret_cast<const float4*>(x), bias,
                                                                 reinterpret_cast<float4*>(y), n / 4, inner / 4, channels,
                                                                 slope, scale);
  else
    bias_act_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(x, bias, y, n, inner, channels, slope, scale);
  return check_launch("bias_act");
}

extern "C" int e4s_noise_bias_act_nhwc_f32(float* x, int batch, int h, int w, int c, const float* noise,
                                           const float* noise_w, int64_t noise_sb, int64_t noise_sc, const float* bias,
                                           float slope, float scale, void* stream) {
  E4S_REQUIRE(x && batch > 0 && h > 0 && w > 0 && c > 0, "noise_bias_act: bad args");
  E4S_REQUIRE(!noise || noise_w, "noise_bias_act: noise without weight");
  int64_t total = (int64_t)batch * h * w * c;
  noise_bias_act_nhwc_kernel<<<grid_for(total, 256, 4), 256, 0, as_stream(stream)>>>(x, total, h * w, w, c, noise, noise_w,
                                                                                     noise_sb, noise_sc, bias, slope, scale);
  return check_launch("noise_bias_act");
}

extern "C" int e4s_mask_labels(const float* mask, int batch, int k, int h, int w, uint8_t* labels, int32_t* flags,
                               void* stream) {
  E4S_REQUIRE(mask && labels && flags && batch > 0 && k > 0 && k < 255 && h > 0 && w > 0, "mask_labels: bad args");
  int64_t hw = (int64_t)h * w, total = hw * batch;
  mask_labels_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(mask, k, hw, total, labels, flags);
  return check_launch("mask_labels");
}

extern "C" int e4s_labels_to_onehot_f32(const uint8_t* labels, int batch, int k, int h, int w, float* onehot, void* stream) {
  E4S_REQUIRE(labels && onehot && batch > 0 && k > 0 && h > 0 && w > 0, "labels_to_onehot: bad args");
  int64_t hw = (int64_t)h * w, total = hw * batch * k;
  labels_to_onehot_kernel<<<grid_for(total, 256, 2), 256, 0, as_stream(stream)>>>(labels, k, hw, total, onehot);
  return check_launch("labels_to_onehot");
}

extern "C" int e4s_torgb_f32(const float* x, int64_t x_pitch, int batch, int h, int w, int cin, const float* smod, const float* wrgb,
                             const uint8_t* labels, int regions, int lab_h, int lab_w, const float* pixw, int64_t pixw_sb,
                             const float* bias, const float* skip, const float* fir, float* rgb, int accumulate,
                             void* stream) {
  E4S_REQUIRE(x && smod && wrgb && rgb && batch > 0 && h > 0 && w > 0, "torgb: bad args");
  E4S_REQUIRE(cin >= 8 && cin % 4 == 0 && x_pitch % 4 == 0, "torgb: cin/x_pitch must be multiples of 4");
  E4S_REQUIRE(!skip || (fir && h % 2 == 0 && w % 2 == 0), "torgb: skip needs fir and even size");
  E4S_REQUIRE(regions > 0 && (!(labels || pixw) || (lab_h > 0 && lab_w > 0)), "torgb: bad region args");
  int64_t npix = (int64_t)batch * h * w;
  cudaStream_t s = as_stream(stream);
  // small images (the 4^2 .. 64^2 layers): 32-pixel tiles, so a CTA's 8 warps share 32 pixels instead of walking 256 one after
  // the other (the kernel is a chain of dependent L2 round trips per pixel step: 75-90 us per launch whatever the size before)
  const bool small = npix <= 32 * 148 * 16;
  const unsigned tiles = (unsigned)ceil_div64(npix, small ? 32 : 256);
  // modulated-weight table in shared memory (see torgb_kernel): large images whose 256-pixel tiles never straddle two samples
  const bool tab = !small && cin >= 128 && cin <= TORGB_TAB_C && regions <= TORGB_TAB_R && ((int64_t)h * w) % 256 == 0 && cin % 4 == 0;
#define E4S_TORGB_LAUNCH(LPV)                                                                                                        \
  do {                                                                                                                             \
    if (small)                                                                                                                     \
      torgb_kernel<LPV, 32><<<tiles, 256, 0, s>>>(x, x_pitch, npix, h, w, cin, smod, wrgb, labels, regions, lab_h, lab_w, pixw, pixw_sb, \
                                                  bias, skip, fir, rgb, accumulate);                                                \
    else if (tab)                                                                                                                  \
      torgb_kernel<LPV, 256, true><<<tiles, 256, 0, s>>>(x, x_pitch, npix, h, w, cin, smod, wrgb, labels, regions, lab_h, lab_w, pixw,  \
                                                         pixw_sb, bias, skip, fir, rgb, accumulate);                                  \
    else                                                                                                                           \
      torgb_kernel<LPV, 256><<<tiles, 256, 0, s>>>(x, x_pitch, npix, h, w, cin, smod, wrgb, labels, regions, lab_h, lab_w, pixw, pixw_sb, \
                                                   bias, skip, fir, rgb, accumulate);                                               \
  } while (0)
  if (cin >= 128) {
    E4S_TORGB_LAUNCH(32);
  } else if (cin >= 32) {
    E4S_TORGB_LAUNCH(8);
  } else {
    E4S_TORGB_LAUNCH(4);
  }
#undef E4S_TORGB_LAUNCH
  return check_launch("torgb");
}

// ---- style recombination between encoder and generator (reference swap_face_fine/swap_face_mask.py:336-367) --------------
// One CTA per (sample, component).  out = target's vector, replaced by the source's where the component is swapped; ears (7)
// are always the average, ear-rings (11) always the target's, below-face (8) the average when asked, and the mouth (9) falls
// back to the target's when the source has no mouth pixels (its masked mean is the zero vector).
namespace e4s {
__global__ void __launch_bounds__(128) swap_comp_styles_kernel(const float* __restrict__ t, const float* __restrict__ s, float* __restrict__ out,
                                                               int ncomp, int dim, uint32_t comp_mask, int below_face) {
  const int c = blockIdx.x, b = blockIdx.y;
  const float* tv = t + ((int64_t)b * ncomp + c) * dim;
  const float* sv = s + ((int64_t)b * ncomp + c) * dim;
  float* ov = out + ((int64_t)b * ncomp + c) * dim;
  int mode = (comp_mask >> c) & 1u;                    // 0 target, 1 source, 2 average
  if (c == 7) mode = 2;
  if (c == 11) mode = 0;
  if (c == 8 && below_face) mode = 2;
  if (c == 9) {                                        // torch.sum(style_vectors2[:, 9, :]) == 0, per sample
    __shared__ float red[4];
    float acc = 0.f;
    for (int i = threadIdx.x; i < dim; i += 128) acc += sv[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (red[0] + red[1] + red[2] + red[3] == 0.f) mode = 0;
  }
  for (int i = threadIdx.x; i < dim; i += 128) {
    const float a = tv[i], bb = sv[i];
    ov[i] = mode == 0 ? a : (mode == 1 ? bb : (a + bb) / 2.f);
  }
}
}  // namespace e4s

extern "C" int e4s_swap_comp_styles_f32(const float* target, const float* source, float* out, int batch, int ncomp, int dim,
                                        uint32_t comp_mask, int below_face, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(target && source && out && batch > 0 && dim > 0, "swap_comp_styles: bad args");
  E4S_REQUIRE(ncomp >= 12 && ncomp <= 32, "swap_comp_styles: the recombination rules address components 7, 8, 9, 11 (ncomp=%d)", ncomp);
  swap_comp_styles_kernel<<<dim3((unsigned)ncomp, (unsigned)batch), 128, 0, as_stream(stream)>>>(target, source, out, ncomp, dim, comp_mask,
                                                                                              below_face);
  return check_launch("swap_comp_styles");
}

// ---- tensor2im on the device (reference utils/torch_utils.py:64-76), batched: [B,3,H,W] float -> [B,H,W,3] uint8 --------------
// (v + 1) / 2 (when zero-centred), clamp to [0,1], * 255, truncate -- the reference's float32 numpy arithmetic, so the bytes are
// identical; a quarter of the D2H traffic of the float image.  Four pixels per thread: float4 loads per plane, three u32 stores.
namespace e4s {
__device__ __forceinline__ uint32_t im_byte(float v, int zero_center) {
  if (zero_center) v = (v + 1.f) / 2.f;
  v = v < 0.f ? 0.f : v;                 // var[var < 0] = 0 (NaN stays NaN in numpy and casts to 0 here as there)
  v = v > 1.f ? 1.f : v;
  return (uint32_t)(int)(v * 255.f);
}
__global__ void __launch_bounds__(256) tensor2im_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int64_t hw4, int64_t hw, int zero_center,
                                                        int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw4, q = i - b * hw4;
    const float* xb = x + b * 3 * hw + q * 4;
    const float4 r = __ldg(reinterpret_cast<const float4*>(xb)), g = __ldg(reinterpret_cast<const float4*>(xb + hw)),
                 bl = __ldg(reinterpret_cast<const float4*>(xb + 2 * hw));
    const uint32_t p0[3] = {im_byte(r.x, zero_center), im_byte(g.x, zero_center), im_byte(bl.x, zero_center)};
    const uint32_t p1[3] = {im_byte(r.y, zero_center), im_byte(g.y, zero_center), im_byte(bl.y, zero_center)};
    const uint32_t p2[3] = {im_byte(r.z, zero_center), im_byte(g.z, zero_center), im_byte(bl.z, zero_center)};
    const uint32_t p3[3] = {im_byte(r.w, zero_center), im_byte(g.w, zero_center), im_byte(bl.w, zero_center)};
    uint32_t* o = reinterpret_cast<uint32_t*>(y + (b * hw + q * 4) * 3);
    o[0] = p0[0] | (p0[1] << 8) | (p0[2] << 16) | (p1[0] << 24);
    o[1] = p1[1] | (p1[2] << 8) | (p2[0] << 16) | (p2[1] << 24);
    o[2] = p2[2] | (p3[0] << 8) | (p3[1] << 16) | (p3[2] << 24);
  }
}
}  // namespace e4s

extern "C" int e4s_tensor2im_u8(const float* x, uint8_t* y, int batch, int h, int w, int zero_center, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(x && y && batch > 0 && h > 0 && w > 0, "tensor2im: bad args");
  E4S_REQUIRE(((int64_t)h * w) % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0,
              "tensor2im: h*w must be a multiple of 4 and the buffers 16 / 4 byte aligned");
  const int64_t hw = (int64_t)h * w, total = (int64_t)batch * (hw / 4);
  tensor2im_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, y, hw / 4, hw, zero_center, total);
  return check_launch("tensor2im");
}

// ---- TO_TENSOR + NORMALIZE on the device (reference datasets/dataset.py:45, torchvision ToTensor / Normalize), batched ------------
// u8 HWC [B,H,W,3] -> fp32 NCHW: x01 = v / 255 (what FaceParser.preprocess_img feeds the parser) and / or
// xn = (x01 - mean[c]) / std[c] (what the pipelines feed Net3) -- the same fp32 operations in the same order, so the floats are
// identical to the reference's.  Four pixels per thread: three u32 loads, one float4 store per plane and output.
namespace e4s {
__global__ void __launch_bounds__(256) im2tensor_kernel(const uint8_t* __restrict__ x, float* __restrict__ y01, float* __restrict__ yn, int64_t hw4,
                                                        int64_t hw, float m0, float m1, float m2, float s0, float s1, float s2, int64_t total) {
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw4, q = i - b * hw4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + (b * hw + q * 4) * 3);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                             (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 v;
      v.x = __fdiv_rn((float)by[c], 255.f); v.y = __fdiv_rn((float)by[3 + c], 255.f);
      v.z = __fdiv_rn((float)by[6 + c], 255.f); v.w = __fdiv_rn((float)by[9 + c], 255.f);
      const int64_t o = (b * 3 + c) * hw + q * 4;
      if (y01) *reinterpret_cast<float4*>(y01 + o) = v;
      if (yn) {
        float4 n;
        n.x = __fdiv_rn(v.x - mean[c], sd[c]); n.y = __fdiv_rn(v.y - mean[c], sd[c]);
        n.z = __fdiv_rn(v.z - mean[c], sd[c]); n.w = __fdiv_rn(v.w - mean[c], sd[c]);
        *reinterpret_cast<float4*>(yn + o) = n;
      }
    }
  }
}
}  // namespace e4s

extern "C" int e4s_im2tensor_f32(const uint8_t* x, float* y01, float* ynorm, int batch, int h, int w, const float* mean3, const float* std3,
                                 void* stream) {
  using namespace e4s;
  E4S_REQUIRE(x && (y01 || ynorm) && batch > 0 && h > 0 && w > 0, "im2tensor: bad args");
  E4S_REQUIRE(!ynorm || (mean3 && std3), "im2tensor: ynorm needs mean3 / std3 (host arrays of 3 floats)");
  E4S_REQUIRE(((int64_t)h * w) % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y01) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ynorm) & 15) == 0,
              "im2tensor: h*w must be a multiple of 4 and the buffers 4 / 16 byte aligned");
  const int64_t hw = (int64_t)h * w, total = (int64_t)batch * (hw / 4);
  const float m[3] = {mean3 ? mean3[0] : 0.f, mean3 ? mean3[1] : 0.f, mean3 ? mean3[2] : 0.f};
  const float sd[3] = {std3 ? std3[0] : 1.f, std3 ? std3[1] : 1.f, std3 ? std3[2] : 1.f};
  im2tensor_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, y01, ynorm, hw / 4, hw, m[0], m[1], m[2], sd[0], sd[1], sd[2], total);
  return check_launch("im2tensor");
}

// ---- grey-scale morphology of the paste-back masks (reference utils/morphology.py:23-200, kornia-style) ----------------------
// out[y,x] = max_{i,j} ( P[y+i, x+j] + nb[se_h-1-i][se_w-1-j] )   (dilation; nb flipped, :93-95)
//          = min_{i,j} ( P[y+i, x+j] - nb[i][j] )                   (erosion, :184-186)
// P = the image padded by the structuring element's origin with `border` (geodesic: -/+ max_val), nb = 0 (or the non-flat
// element) where kernel != 0 and -max_val elsewhere.  One thread per output pixel, the element cached in shared memory.
namespace e4s {
template <bool DILATE>
__global__ void __launch_bounds__(256) morphology_kernel(const float* __restrict__ x, const float* __restrict__ nb, float* __restrict__ out,
                                                         int h, int w, int se_h, int se_w, int oy, int ox, float border, int64_t total) {
  extern __shared__ float s_nb[];
  for (int i = threadIdx.x; i < se_h * se_w; i += blockDim.x) s_nb[i] = nb[i];
  __syncthreads();
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int px = (int)(idx % w);
    const int64_t t = idx / w;
    const int py = (int)(t % h);
    const float* plane = x + (t / h) * (int64_t)h * w;
    float acc = DILATE ? -INFINITY : INFINITY;
    for (int i = 0; i < se_h; ++i) {
      const int iy = py + i - oy;
      const bool rowin = iy >= 0 && iy < h;
      for (int j = 0; j < se_w; ++j) {
        const int ix = px + j - ox;
        const float v = (rowin && ix >= 0 && ix < w) ? __ldg(plane + (int64_t)iy * w + ix) : border;
        if (DILATE) acc = fmaxf(acc, v + s_nb[(se_h - 1 - i) * se_w + (se_w - 1 - j)]);
        else acc = fminf(acc, v - s_nb[i * se_w + j]);
      }
    }
    out[idx] = acc;
  }
}
}  // namespace e4s

extern "C" int e4s_morphology_f32(const float* x, const float* neighborhood, float* out, int64_t planes, int h, int w, int se_h, int se_w,
                                  int origin_y, int origin_x, float border_value, int dilate, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(x && neighborhood && out && planes > 0 && h > 0 && w > 0, "morphology: bad args");
  E4S_REQUIRE(se_h > 0 && se_w > 0 && se_h * se_w <= 8192 && origin_y >= 0 && origin_y < se_h && origin_x >= 0 && origin_x < se_w,
              "morphology: bad structuring element %dx%d origin (%d,%d)", se_h, se_w, origin_y, origin_x);
  const int64_t total = planes * h * w;
  const size_t smem = (size_t)se_h * se_w * sizeof(float);
  if (dilate)
    morphology_kernel<true><<<grid_for(total, 256), 256, smem, as_stream(stream)>>>(x, neighborhood, out, h, w, se_h, se_w, origin_y, origin_x,
                                                                                  border_value, total);
  else
    morphology_kernel<false><<<grid_for(total, 256), 256, smem, as_stream(stream)>>>(x, neighborhood, out, h, w, se_h, se_w, origin_y, origin_x,
                                                                                   border_value, total);
  return check_launch("morphology");
}

// ---- gradient of the fused bias + leaky ReLU (reference fused_bias_act_kernel.cu:26-47, act = 3, grad = 1) ---------------------
// y = ((g + bias?) taken with slope where the forward OUTPUT ref was <= 0) * scale; used for grad_input (bias = NULL) and for the
// second-order term (reference op/fused_act.py:18-47).
namespace e4s {
__global__ void bias_act_grad_kernel(const float* __restrict__ g, const float* __restrict__ bias, const float* __restrict__ ref,
                                     float* __restrict__ y, int64_t n, int64_t inner, int channels, float slope, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    if (bias) v += __ldg(bias + (i / inner) % channels);
    y[i] = (ref[i] > 0.f ? v : v * slope) * scale;
  }
}
}  // namespace e4s

extern "C" int e4s_bias_act_grad_f32(const float* g, const float* bias, const float* ref, float* y, int64_t n, int64_t inner, int channels,
                                     float slope, float scale, void* stream) {
  using namespace e4s;
  E4S_REQUIRE(g && ref && y && n > 0 && inner > 0 && channels > 0, "bias_act_grad: bad args");
  bias_act_grad_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(g, bias, ref, y, n, inner, channels, slope, scale);
  return check_launch("bias_act_grad");
}


// Backward-pass kernels of the regional modulated convolution (PTI fine-tuning runs loss.backward() through
// net.G, reference training/video_swap_ft_coach.py:242-318; the reference differentiates F.conv2d / F.conv_transpose2d through
// models/stylegan2/op/conv2d_gradfix.py:134-225).  The data gradient re-uses the forward engines with transposed weights
// (e4s_conv_tc / e4s_conv_f32); what has no forward counterpart lives here:
//   e4s_conv_wgrad_f32      weight gradient: a [cout x cin] GEMM per tap whose K dimension is the pixels (fp32 CUDA cores, split over
//                           pixel chunks with a fixed-order second pass: deterministic)
//   e4s_region_scale_f32    g[p,c] * table[b, r(p), c], optionally keeping only the pixels of one region (the reference's mask multiply)
//   e4s_region_dot_f32      out[b,r,c] = sum over the pixels of region r of a[p,c] * b[p,c] (gradients of the style / demodulation tables)
//   e4s_chan_scale_accum_f32  dx[p,c] (+)= h[p,c] * s[b,c]
#include "common.cuh"

namespace e4s {

// ---------------------------------------------------------------------------------------------------------------------------------
// dW[co][ci][ky][kx] (+)= sum_b sum_{(y,x) < (hl,wl)} sscale[b,ci] * X[b, y + ky*tx - px, x + kx*tx - px, ci] * G[b, y*sg + ky*tg - pg, x*sg + kx*tg - pg, co]
//   same-resolution 3x3 (pad 1): loop over output pixels, tx = 1, px = 1, sg = 1, tg = 0, pg = 0
//   conv_transpose2d(stride 2):   loop over input pixels,  tx = 0, px = 0, sg = 2, tg = 1, pg = 0   (G on the (2H+1) x (2W+1) grid)
//   linear layers:                kh = kw = 1, hl = rows, wl = 1
// One CTA = a 64 (co) x 64 (ci) block of one tap over one chunk of loop pixels; 256 threads, 4 x 4 outputs each.
// ---------------------------------------------------------------------------------------------------------------------------------
struct WgradGeom {
  int batch, hx, wx, cin, hg, wg, cout, hl, wl, kh, kw, tx, px, sg, tg, pg;
  int64_t x_pitch, g_pitch, s_stride;
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ sscale,
                                                        const WgradGeom q, const int nchunks, const int64_t chunk_px, float* __restrict__ partial) {
  __shared__ __align__(16) float sx[16][64];
  __shared__ __align__(16) float sg_[16][64];
  const int co0 = (int)blockIdx.x * 64, ci0 = (int)blockIdx.y * 64;
  const int tap = (int)blockIdx.z / nchunks, chunk = (int)blockIdx.z % nchunks;
  const int ky = tap / q.kw, kx = tap - ky * q.kw;
  const int64_t total = (int64_t)q.batch * q.hl * q.wl;
  const int64_t p0 = (int64_t)chunk * chunk_px, p1 = p0 + chunk_px < total ? p0 + chunk_px : total;
  const int tid = threadIdx.x;
  const int lk = tid >> 4, lc = (tid & 15) * 4;          // loader: pixel lk of the stage, channels lc .. lc+3
  const int tco = (tid >> 4) * 4, tci = (tid & 15) * 4;  // compute: 4 x 4 block
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t pb = p0; pb < p1; pb += 16) {
    const int64_t pidx = pb + lk;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pidx < p1) {
      const int xw = (int)(pidx % q.wl);
      const int64_t t = pidx / q.wl;
      const int yh = (int)(t % q.hl), b = (int)(t / q.hl);
      const int xy = yh + ky * q.tx - q.px, xx = xw + kx * q.tx - q.px;
      const int gy = yh * q.sg + ky * q.tg - q.pg, gx = xw * q.sg + kx * q.tg - q.pg;
      if (xy >= 0 && xy < q.hx && xx >= 0 && xx < q.wx && gy >= 0 && gy < q.hg && gx >= 0 && gx < q.wg) {
        const int ci = ci0 + lc, co = co0 + lc;
        if (ci < q.cin) {
          xv = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * q.hx + xy) * q.wx + xx) * q.x_pitch + ci));
          if (ci + 1 >= q.cin) xv.y = 0.f;
          if (ci + 2 >= q.cin) xv.z = 0.f;
          if (ci + 3 >= q.cin) xv.w = 0.f;
          if (sscale) {
            const float* sp = sscale + (int64_t)b * q.s_stride + ci;
            xv.x *= __ldg(sp);
            if (ci + 1 < q.cin) xv.y *= __ldg(sp + 1);
            if (ci + 2 < q.cin) xv.z *= __ldg(sp + 2);
            if (ci + 3 < q.cin) xv.w *= __ldg(sp + 3);
          }
        }
        if (co < q.cout) {
          gv = __ldg(reinterpret_cast<const float4*>(g + (((int64_t)b * q.hg + gy) * q.wg + gx) * q.g_pitch + co));
          if (co + 1 >= q.cout) gv.y = 0.f;
          if (co + 2 >= q.cout) gv.z = 0.f;
          if (co + 3 >= q.cout) gv.w = 0.f;
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&sx[lk][lc]) = xv;
    *reinterpret_cast<float4*>(&sg_[lk][lc]) = gv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sg_[k][tco]);
      const float4 bb = *reinterpret_cast<const float4*>(&sx[k][tci]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  // partial [chunk][tap][co][ci]
  const int taps = q.kh * q.kw;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tco + i;
    if (co >= q.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tci + j;
      if (ci < q.cin) partial[(((int64_t)chunk * taps + tap) * q.cout + co) * q.cin + ci] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] (+)= scale * sum_chunks partial[chunk][tap][co][ci]   (fixed order)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int taps, int cout, int cin, float scale,
                                                          int accumulate, float* __restrict__ dw) {
  const int64_t n = (int64_t)taps * cout * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += partial[(int64_t)c * n + i];
    const int ci = (int)(i % cin);
    const int64_t t = i / cin;
    const int co = (int)(t % cout), tap = (int)(t / cout);
    float* o = dw + ((int64_t)co * cin + ci) * taps + tap;
    *o = accumulate ? *o + s * scale : s * scale;
  }
}

__global__ void __launch_bounds__(256) region_scale_kernel(const float* __restrict__ g, int64_t g_pitch, int batch, int h, int w, int c,
                                                          const float* __restrict__ table, const uint8_t* __restrict__ labels, int regions, int lab_h,
                                                          int lab_w, int select, float* __restrict__ out, int64_t out_pitch, int out_c) {
  // one thread per (pixel, float4 of the out_c output channels); channels >= c of the output are zero (padding for the engines' cin % 8)
  const int c4 = out_c >> 2;
  const int64_t total = (int64_t)batch * h * w * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    const int64_t pix = i / c4;
    const int x = (int)(pix % w);
    const int64_t t = pix / w;
    const int y = (int)(t % h), b = (int)(t / h);
    int r = 0;
    if (labels) r = labels[((int64_t)b * lab_h + nearest_src(y, lab_h, h)) * lab_w + nearest_src(x, lab_w, w)];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (select < 0 || r == select) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cc + j < c) {
          v[j] = g[pix * g_pitch + cc + j];
          if (table) v[j] *= __ldg(table + ((int64_t)b * regions + r) * c + cc + j);
        }
      }
    }
    *reinterpret_cast<float4*>(out + pix * out_pitch + cc) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// partial[b][chunk][r][c] = sum over the chunk's pixels of region r of a*b (fp64); chunk = 1024 pixels of one sample (batch-invariant split)
constexpr int RD_CHUNK = 1024;
__global__ void __launch_bounds__(128) region_dot_kernel(const float* __restrict__ a, int64_t a_pitch, const float* __restrict__ bsrc, int64_t b_pitch,
                                                        int h, int w, int c, const uint8_t* __restrict__ labels, int regions, int lab_h, int lab_w,
                                                        int nchunks, double* __restrict__ partial) {
  extern __shared__ double racc[];                       // [regions][128]
  const int chunk = (int)blockIdx.x % nchunks, b = (int)blockIdx.x / nchunks;
  const int ch = (int)blockIdx.y * 128 + threadIdx.x;
  for (int r = 0; r < regions; ++r) racc[r * 128 + threadIdx.x] = 0.0;
  const int hw = h * w;
  const int p0 = chunk * RD_CHUNK, p1 = min(p0 + RD_CHUNK, hw);
  double cur = 0.0;
  int cur_r = -1;
  for (int p = p0; p < p1; ++p) {
    int r = 0;
    if (labels) {
      const int y = p / w, x = p - y * w;
      r = labels[((int64_t)b * lab_h + nearest_src(y, lab_h, h)) * lab_w + nearest_src(x, lab_w, w)];
      if (r >= regions) r = regions - 1;
    }
    if (r != cur_r) {
      if (cur_r >= 0) racc[cur_r * 128 + threadIdx.x] += cur;
      cur = 0.0;
      cur_r = r;
    }
    if (ch < c) {
      const int64_t pix = (int64_t)b * hw + p;
      cur += (double)a[pix * a_pitch + ch] * (double)bsrc[pix * b_pitch + ch];
    }
  }
  if (cur_r >= 0) racc[cur_r * 128 + threadIdx.x] += cur;
  if (ch < c)
    for (int r = 0; r < regions; ++r) partial[(((int64_t)b * nchunks + chunk) * regions + r) * c + ch] = racc[r * 128 + threadIdx.x];
}

__global__ void __launch_bounds__(256) region_dot_reduce_kernel(const double* __restrict__ partial, int batch, int nchunks, int regions, int c,
                                                               float* __restrict__ out) {
  const int64_t n = (int64_t)batch * regions * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rc = i % ((int64_t)regions * c);
    const int b = (int)(i / ((int64_t)regions * c));
    double s = 0.0;
    for (int k = 0; k < nchunks; ++k) s += partial[((int64_t)b * nchunks + k) * regions * c + rc];
    out[i] = (float)s;
  }
}

__global__ void __launch_bounds__(256) chan_scale_accum_kernel(const float* __restrict__ hsrc, int64_t h_pitch, const float* __restrict__ s,
                                                              int64_t s_stride, float* __restrict__ dx, int64_t dx_pitch, int batch, int64_t hw, int c,
                                                              int accumulate) {
  const int c4 = c >> 2;
  const int64_t total = (int64_t)batch * hw * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    const int64_t pix = i / c4;
    const int b = (int)(pix / hw);
    const float4 hv = *reinterpret_cast<const float4*>(hsrc + pix * h_pitch + cc);
    float4 sv = make_float4(1.f, 1.f, 1.f, 1.f);
    if (s) sv = __ldg(reinterpret_cast<const float4*>(s + (int64_t)b * s_stride + cc));
    float4* o = reinterpret_cast<float4*>(dx + pix * dx_pitch + cc);
    float4 v = make_float4(hv.x * sv.x, hv.y * sv.y, hv.z * sv.z, hv.w * sv.w);
    if (accumulate) {
      const float4 old = *o;
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
    }
    *o = v;
  }
}

static inline unsigned ew_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  return (unsigned)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

static int wgrad_chunks(const WgradGeom& q) {
  const int64_t base = (int64_t)ceil_div(q.cout, 64) * ceil_div(q.cin, 64) * q.kh * q.kw;
  const int64_t total = (int64_t)q.batch * q.hl * q.wl;
  int64_t n = ceil_div64(4 * 148, base);
  const int64_t maxn = ceil_div64(total, 256);            // at least 256 pixels per chunk
  if (n > maxn) n = maxn;
  if (n > 256) n = 256;
  if (n < 1) n = 1;
  return (int)n;
}

}  // namespace e4s

using namespace e4s;

extern "C" int64_t e4s_conv_wgrad_ws_bytes(int batch, int hl, int wl, int cin, int cout, int kh, int kw) {
  WgradGeom q{};
  q.batch = batch; q.hl = hl; q.wl = wl; q.cin = cin; q.cout = cout; q.kh = kh; q.kw = kw;
  return (int64_t)wgrad_chunks(q) * kh * kw * cin * cout * 4;
}

extern "C" int e4s_conv_wgrad_f32(const float* x, int64_t x_pitch, int batch, int hx, int wx, int cin, const float* g, int64_t g_pitch, int hg, int wg,
                                  int cout, int hl, int wl, int kh, int kw, int tx, int px, int sg, int tg, int pg, const float* sscale,
                                  int64_t s_stride, float scale, float* dw, int accumulate, void* ws, void* stream) {
  E4S_REQUIRE(x && g && dw && ws && batch > 0 && hx > 0 && wx > 0 && cin > 0 && hg > 0 && wg > 0 && cout > 0 && hl > 0 && wl > 0 && kh > 0 && kw > 0,
              "conv_wgrad: bad args");
  E4S_REQUIRE(x_pitch % 4 == 0 && g_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0,
              "conv_wgrad: x / g must be 16-byte aligned with pitch %% 4 == 0");
  E4S_REQUIRE(x_pitch >= ((cin + 3) & ~3) && g_pitch >= ((cout + 3) & ~3), "conv_wgrad: pitch must cover the channels rounded up to 4");
  WgradGeom q{batch, hx, wx, cin, hg, wg, cout, hl, wl, kh, kw, tx, px, sg, tg, pg, x_pitch, g_pitch, s_stride};
  const int nchunks = wgrad_chunks(q);
  const int64_t total = (int64_t)batch * hl * wl;
  const int64_t chunk_px = ceil_div64(ceil_div64(total, nchunks), 16) * 16;
  const int64_t gz = (int64_t)kh * kw * nchunks;
  E4S_REQUIRE(gz < 65536, "conv_wgrad: too many chunks");
  cudaStream_t s = as_stream(stream);
  conv_wgrad_kernel<<<dim3((unsigned)ceil_div(cout, 64), (unsigned)ceil_div(cin, 64), (unsigned)gz), 256, 0, s>>>(x, g, sscale, q, nchunks, chunk_px,
                                                                                                           static_cast<float*>(ws));
  int rc = check_launch("conv_wgrad");
  if (rc) return rc;
  wgrad_reduce_kernel<<<ew_grid((int64_t)kh * kw * cin * cout), 256, 0, s>>>(static_cast<const float*>(ws), nchunks, kh * kw, cout, cin, scale, accumulate, dw);
  return check_launch("conv_wgrad(reduce)");
}

extern "C" int e4s_region_scale_f32(const float* g, int64_t g_pitch, int batch, int h, int w, int c, const float* table, const uint8_t* labels,
                                    int regions, int lab_h, int lab_w, int select_region, float* out, int64_t out_pitch, int out_c, void* stream) {
  E4S_REQUIRE(g && out && batch > 0 && h > 0 && w > 0 && c > 0 && out_c >= c && out_c % 4 == 0 && out_pitch % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0 && regions >= 1 && (!labels || (lab_h > 0 && lab_w > 0)),
              "region_scale: bad args");
  region_scale_kernel<<<ew_grid((int64_t)batch * h * w * (out_c / 4)), 256, 0, as_stream(stream)>>>(g, g_pitch, batch, h, w, c, table, labels, regions,
                                                                                                 lab_h, lab_w, select_region, out, out_pitch, out_c);
  return check_launch("region_scale");
}

extern "C" int64_t e4s_region_dot_ws_bytes(int batch, int h, int w, int c, int regions) {
  return (int64_t)batch * ceil_div(h * w, RD_CHUNK) * regions * c * 8;
}

extern "C" int e4s_region_dot_f32(const float* a, int64_t a_pitch, const float* b, int64_t b_pitch, int batch, int h, int w, int c,
                                  const uint8_t* labels, int regions, int lab_h, int lab_w, float* out, void* ws, void* stream) {
  E4S_REQUIRE(a && b && out && ws && batch > 0 && h > 0 && w > 0 && c > 0 && regions >= 1 && regions <= 32 && (!labels || (lab_h > 0 && lab_w > 0)),
              "region_dot: bad args");
  const int nchunks = ceil_div(h * w, RD_CHUNK);
  cudaStream_t s = as_stream(stream);
  region_dot_kernel<<<dim3((unsigned)(batch * nchunks), (unsigned)ceil_div(c, 128)), 128, (size_t)regions * 128 * sizeof(double), s>>>(
      a, a_pitch, b, b_pitch, h, w, c, labels, regions, lab_h, lab_w, nchunks, static_cast<double*>(ws));
  int rc = check_launch("region_dot");
  if (rc) return rc;
  region_dot_reduce_kernel<<<ew_grid((int64_t)batch * regions * c), 256, 0, s>>>(static_cast<const double*>(ws), batch, nchunks, regions, c, out);
  return check_launch("region_dot(reduce)");
}

extern "C" int e4s_chan_scale_accum_f32(const float* h, int64_t h_pitch, const float* s, int64_t s_stride, float* dx, int64_t dx_pitch, int batch,
                                        int64_t hw, int c, int accumulate, void* stream) {
  E4S_REQUIRE(h && dx && batch > 0 && hw > 0 && c > 0 && c % 4 == 0 && h_pitch % 4 == 0 && dx_pitch % 4 == 0 && s_stride % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (!s || (reinterpret_cast<uintptr_t>(s) & 15) == 0),
              "chan_scale_accum: bad args / alignment");
  chan_scale_accum_kernel<<<ew_grid((int64_t)batch * hw * (c / 4)), 256, 0, as_stream(stream)>>>(h, h_pitch, s, s_stride, dx, dx_pitch, batch, hw, c, accumulate);
  return check_launch("chan_scale_accum");
}

// Resampling kernels: bilinear resize (both torch conventions) fused with the NCHW<->NHWC layout change,
// BiSeNet's x8 align_corners upsample fused with argmax + label LUT (the [B,19,512,512] logits are never
// written), and the FaceParser front-end (separable bicubic down-sampling + clamp + normalise).
#include "common.cuh"

namespace e4s {

// torch area_pixel_compute_source_index (linear modes)
__device__ __forceinline__ float src_index(float scale, int dst, bool align_corners) {
  if (align_corners) return scale * (float)dst;
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}
__host__ inline float resize_scale(int in, int out, bool align_corners) {
  if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp make_lerp(float scale, int dst, int in, bool ac) {
  float s = src_index(scale, dst, ac);
  int i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  Lerp r;
  r.i0 = i0;
  r.i1 = i0 + (i0 < in - 1 ? 1 : 0);
  r.l1 = s - (float)i0;
  r.l0 = 1.f - r.l1;
  return r;
}

__global__ void resize_nchw_to_nhwc_kernel(const float* __restrict__ x, int c, int hin, int win, float* __restrict__ y,
                                           int hout, int wout, int c_pad, float sh, float sw, int ac, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c_pad);
    int64_t t = i / c_pad;
    const int ox = (int)(t % wout);
    t /= wout;
    const int oy = (int)(t % hout);
    const int b = (int)(t / hout);
    float v = 0.f;
    if (ch < c) {
      Lerp ly = make_lerp(sh, oy, hin, ac), lx = make_lerp(sw, ox, win, ac);
      const float* p = x + ((int64_t)b * c + ch) * hin * win;
      v = ly.l0 * (lx.l0 * __ldg(p + (int64_t)ly.i0 * win + lx.i0) + lx.l1 * __ldg(p + (int64_t)ly.i0 * win + lx.i1)) +
          ly.l1 * (lx.l0 * __ldg(p + (int64_t)ly.i1 * win + lx.i0) + lx.l1 * __ldg(p + (int64_t)ly.i1 * win + lx.i1));
    }
    y[i] = v;
  }
}

__global__ void resize_nhwc_to_nchw_kernel(const float* __restrict__ x, int64_t pitch, int c, int hin, int win,
                                           float* __restrict__ y, int hout, int wout, float sh, float sw, int ac,
                                           int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wout);
    int64_t t = i / wout;
    const int oy = (int)(t % hout);
    t /= hout;
    const int ch = (int)(t % c);
    const int b = (int)(t / c);
    Lerp ly = make_lerp(sh, oy, hin, ac), lx = make_lerp(sw, ox, win, ac);
    const float* p = x + (int64_t)b * hin * win * pitch + ch;
    auto at = [&](int yy, int xx) { return __ldg(p + ((int64_t)yy * win + xx) * pitch); };
    y[i] = ly.l0 * (lx.l0 * at(ly.i0, lx.i0) + lx.l1 * at(ly.i0, lx.i1)) +
           ly.l1 * (lx.l0 * at(ly.i1, lx.i0) + lx.l1 * at(ly.i1, lx.i1));
  }
}

// one thread per output pixel; C <= 32 logits, 4-channel vector loads from the (small, L1/L2 resident) source
template <int CMAX>
__global__ void upsample_argmax_kernel(const float* __restrict__ x, int64_t pitch, int c, int hin, int win, int hout,
                                       int wout, float sh, float sw, const uint8_t* __restrict__ lut,
                                       uint8_t* __restrict__ labels, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wout);
    int64_t t = i / wout;
    const int oy = (int)(t % hout);
    const int b = (int)(t / hout);
    Lerp ly = make_lerp(sh, oy, hin, true), lx = make_lerp(sw, ox, win, true);
    const float* base = x + (int64_t)b * hin * win * pitch;
    const float* p00 = base + ((int64_t)ly.i0 * win + lx.i0) * pitch;
    const float* p01 = base + ((int64_t)ly.i0 * win + lx.i1) * pitch;
    const float* p10 = base + ((int64_t)ly.i1 * win + lx.i0) * pitch;
    const float* p11 = base + ((int64_t)ly.i1 * win + lx.i1) * pitch;
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int q = 0; q < CMAX / 4; ++q) {
      if (q * 4 >= c) break;
      float4 a = __ldg(reinterpret_cast<const float4*>(p00) + q), bb = __ldg(reinterpret_cast<const float4*>(p01) + q);
      float4 cc = __ldg(reinterpret_cast<const float4*>(p10) + q), d = __ldg(reinterpret_cast<const float4*>(p11) + q);
      float v[4];
      v[0] = ly.l0 * (lx.l0 * a.x + lx.l1 * bb.x) + ly.l1 * (lx.l0 * cc.x + lx.l1 * d.x);
      v[1] = ly.l0 * (lx.l0 * a.y + lx.l1 * bb.y) + ly.l1 * (lx.l0 * cc.y + lx.l1 * d.y);
      v[2] = ly.l0 * (lx.l0 * a.z + lx.l1 * bb.z) + ly.l1 * (lx.l0 * cc.z + lx.l1 * d.z);
      v[3] = ly.l0 * (lx.l0 * a.w + lx.l1 * bb.w) + ly.l1 * (lx.l0 * cc.w + lx.l1 * d.w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = q * 4 + j;
        if (ch < c && v[j] > best) {  // strict '>' keeps the lowest index on ties, like torch.argmax
          best = v[j];
          arg = ch;
        }
      }
    }
    labels[i] = lut ? lut[arg] : (uint8_t)arg;
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// vertical pass: x [planes, hin, win] -> tmp [planes, hin/factor, win]
__global__ void bicubic_v_kernel(const float* __restrict__ x, int hin, int win, int factor, const float* __restrict__ taps,
                                 float* __restrict__ tmp, int64_t total) {
  const int n = factor * 4, pad0 = (n - factor) / 2, ho = hin / factor;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % win);
    int64_t t = i / win;
    const int oy = (int)(t % ho);
    const int64_t plane = t / ho;
    const float* p = x + plane * hin * win + xx;
    float v = 0.f;
    for (int k = 0; k < n; ++k) v = fmaf(__ldg(p + (int64_t)reflect_idx(oy * factor + k - pad0, hin) * win), __ldg(taps + k), v);
    tmp[i] = v;
  }
}

// horizontal pass + clamp + normalise, writes NHWC (c_pad channels, zero filled)
__global__ void bicubic_h_norm_kernel(const float* __restrict__ tmp, int ho, int win, int factor,
                                      const float* __restrict__ taps, const float* __restrict__ mean,
                                      const float* __restrict__ stdv, float* __restrict__ y, int c_pad, int do_clamp, int64_t total) {
  const int n = factor * 4, pad0 = (n - factor) / 2, wo = win / factor;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c_pad);
    int64_t t = i / c_pad;
    const int ox = (int)(t % wo);
    t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    float v = 0.f;
    if (ch < 3) {
      const float* p = tmp + (((int64_t)b * 3 + ch) * ho + oy) * win;
      for (int k = 0; k < n; ++k) v = fmaf(__ldg(p + reflect_idx(ox * factor + k - pad0, win)), __ldg(taps + k), v);
      if (do_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
      v = (v - __ldg(mean + ch)) / __ldg(stdv + ch);
    }
    y[i] = v;
  }
}

// Both passes in one kernel (factor <= 4): a block produces a 16 x 32 tile of output pixels for the three channels from a shared-memory
// copy of the (16f + 3f) x (32f + 3f) input window (reflect padding applied while loading).  Same arithmetic as the two kernels above --
// vertical fmaf chain over the taps in order, then the horizontal one, clamp, normalise -- so the results are bit-identical; what is
// saved is the [planes, h/f, w] intermediate in HBM (written and read back) and the strided 4-byte NHWC stores: pixels leave as whole
// c_pad-channel rows.
constexpr int BC_TY = 16, BC_TX = 32;
__global__ void __launch_bounds__(256) bicubic_fused_kernel(const float* __restrict__ x, int hin, int win, int factor, const float* __restrict__ taps,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ y,
                                                           int c_pad, int do_clamp) {
  extern __shared__ float bsm[];
  const int n = factor * 4, pad0 = (n - factor) / 2, ho = hin / factor, wo = win / factor;
  const int rows = BC_TY * factor + n - factor, cols = BC_TX * factor + n - factor;
  float* sin_ = bsm;                                  // [rows][cols]
  float* sv = sin_ + rows * cols;                     // [BC_TY][cols]
  float* so = sv + BC_TY * cols;                      // [BC_TY][BC_TX][4]: channels 0..2 (+ one zero)
  float* stap = so + BC_TY * BC_TX * 4;               // [n]
  const int b = blockIdx.z, oy0 = blockIdx.y * BC_TY, ox0 = blockIdx.x * BC_TX;
  const int tid = threadIdx.x;
  if (tid < n) stap[tid] = __ldg(taps + tid);
  for (int ch = 0; ch < 3; ++ch) {
    const float* xp = x + ((int64_t)b * 3 + ch) * hin * win;
    __syncthreads();                                  // previous channel's readers are done (also publishes stap)
    for (int i = tid; i < rows * cols; i += 256) {
      const int r = i / cols, c = i - r * cols;
      // partial tiles at the bottom / right edge read rows past the reflection range: clamp (those outputs are never stored)
      const int iy = min(max(reflect_idx(oy0 * factor + r - pad0, hin), 0), hin - 1), ix = min(max(reflect_idx(ox0 * factor + c - pad0, win), 0), win - 1);
      sin_[i] = __ldg(xp + (int64_t)iy * win + ix);
    }
    __syncthreads();
    for (int i = tid; i < BC_TY * cols; i += 256) {
      const int ty = i / cols, c = i - ty * cols;
      float v = 0.f;
      for (int k = 0; k < n; ++k) v = fmaf(sin_[(ty * factor + k) * cols + c], stap[k], v);
      sv[i] = v;
    }
    __syncthreads();
    const float m = __ldg(mean + ch), sd = __ldg(stdv + ch);
    for (int i = tid; i < BC_TY * BC_TX; i += 256) {
      const int ty = i / BC_TX, tx = i - ty * BC_TX;
      float v = 0.f;
      for (int k = 0; k < n; ++k) v = fmaf(sv[ty * cols + tx * factor + k], stap[k], v);
      if (do_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
      so[i * 4 + ch] = (v - m) / sd;
    }
  }
  __syncthreads();
  // NHWC rows of c_pad floats (c_pad % 4 == 0): first float4 = (c0, c1, c2, 0), the rest zero
  const int q4 = c_pad >> 2;
  for (int i = tid; i < BC_TY * BC_TX * q4; i += 256) {
    const int q = i % q4, px = i / q4;
    const int ty = px / BC_TX, tx = px - ty * BC_TX;
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy >= ho || ox >= wo) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q == 0) v = make_float4(so[px * 4], so[px * 4 + 1], so[px * 4 + 2], 0.f);
    *reinterpret_cast<float4*>(y + (((int64_t)b * ho + oy) * wo + ox) * c_pad + 4 * q) = v;
  }
}

static inline unsigned grid_for(int64_t n, int block) {
  int64_t g = ceil_div64(n, block);
  const int64_t cap = 148 * 64;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace e4s

using namespace e4s;

extern "C" int e4s_resize_bilinear_nchw_to_nhwc_f32(const float* x, int batch, int c, int hin, int win, float* y, int hout,
                                                    int wout, int c_pad, int align_corners, void* stream) {
  E4S_REQUIRE(x && y && batch > 0 && c > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0 && c_pad >= c, "resize: bad args");
  int64_t total = (int64_t)batch * hout * wout * c_pad;
  resize_nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      x, c, hin, win, y, hout, wout, c_pad, resize_scale(hin, hout, align_corners), resize_scale(win, wout, align_corners),
      align_corners, total);
  return check_launch("resize_nchw_to_nhwc");
}

extern "C" int e4s_resize_bilinear_nhwc_to_nchw_f32(const float* x, int64_t x_pitch, int batch, int c, int hin, int win,
                                                    float* y, int hout, int wout, int align_corners, void* stream) {
  E4S_REQUIRE(x && y && batch > 0 && c > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0 && x_pitch >= c, "resize: bad args");
  int64_t total = (int64_t)batch * hout * wout * c;
  resize_nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      x, x_pitch, c, hin, win, y, hout, wout, resize_scale(hin, hout, align_corners), resize_scale(win, wout, align_corners),
      align_corners, total);
  return check_launch("resize_nhwc_to_nchw");
}

extern "C" int e4s_upsample_argmax_u8(const float* logits, int64_t pitch, int batch, int c, int hin, int win, int hout,
                                      int wout, const uint8_t* lut, uint8_t* labels, void* stream) {
  E4S_REQUIRE(logits && labels && batch > 0 && c > 0 && c <= 32 && hin > 0 && win > 0 && hout > 0 && wout > 0,
              "upsample_argmax: bad args (c must be <= 32)");
  E4S_REQUIRE(pitch % 4 == 0 && pitch >= ((c + 3) / 4) * 4 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0,
              "upsample_argmax: pitch must be a multiple of 4 covering c");
  int64_t total = (int64_t)batch * hout * wout;
  upsample_argmax_kernel<32><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      logits, pitch, c, hin, win, hout, wout, resize_scale(hin, hout, true), resize_scale(win, wout, true), lut, labels, total);
  return check_launch("upsample_argmax");
}

extern "C" int e4s_bicubic_down_norm_f32(const float* x, int batch, int hin, int win, int factor, const float* taps,
                                         const float* mean, const float* stdv, float* tmp, float* y, int c_pad, int do_clamp, void* stream) {
  E4S_REQUIRE(x && taps && mean && stdv && tmp && y, "bicubic: null pointer");
  E4S_REQUIRE(batch > 0 && factor >= 1 && hin % factor == 0 && win % factor == 0 && c_pad >= 3, "bicubic: bad shape");
  E4S_REQUIRE(hin > 4 * factor && win > 4 * factor, "bicubic: image too small for reflect padding");
  const int ho = hin / factor, wo = win / factor;
  if (factor <= 4 && c_pad % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const int n = factor * 4, rows = BC_TY * factor + n - factor, cols = BC_TX * factor + n - factor;
    const size_t smem = (size_t)(rows * cols + BC_TY * cols + BC_TY * BC_TX * 4 + n) * sizeof(float);
    static bool attr_dev[E4S_MAX_DEVICES] = {};
    bool& attr = attr_dev[current_device_slot()];
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(bicubic_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "bicubic: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr = true;
    }
    bicubic_fused_kernel<<<dim3((unsigned)ceil_div(wo, BC_TX), (unsigned)ceil_div(ho, BC_TY), (unsigned)batch), 256, smem, as_stream(stream)>>>(
        x, hin, win, factor, taps, mean, stdv, y, c_pad, do_clamp);
    return check_launch("bicubic_fused");
  }
  int64_t t1 = (int64_t)batch * 3 * ho * win;
  bicubic_v_kernel<<<grid_for(t1, 256), 256, 0, as_stream(stream)>>>(x, hin, win, factor, taps, tmp, t1);
  int rc = check_launch("bicubic_v");
  if (rc) return rc;
  int64_t t2 = (int64_t)batch * ho * wo * c_pad;
  bicubic_h_norm_kernel<<<grid_for(t2, 256), 256, 0, as_stream(stream)>>>(tmp, ho, win, factor, taps, mean, stdv, y, c_pad, do_clamp, t2);
  return check_launch("bicubic_h_norm");
}

// Paste-back mask / blending kernels: SoftErosion (reference utils/paste_back_tricks.py:17-42) and the
// Laplacian-pyramid blend built from cv2.pyrDown / cv2.pyrUp arithmetic (swap_face_fine/multi_band_blending.py:6-74).  All operate on
// NCHW fp32 planes [planes, h, w]; HBM / L2-bound stencils, one output pixel per thread.
#include "common.cuh"

namespace e4s {

static inline unsigned pb_grid(int64_t n, int block = 256) {
  int64_t g = ceil_div64(n, block);
  return (unsigned)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

// depthwise k x k correlation with zero padding k/2 (F.conv2d(x, weight, groups=C, padding=r)); min_with_input: out = min(x, conv(x))
constexpr int DW_TILE = 32;
__global__ void __launch_bounds__(DW_TILE * 8) depthwise_conv_kernel(const float* __restrict__ x, const float* __restrict__ wt, float* __restrict__ out, int h, int w,
                                                                     int k, int min_with_input) {
  extern __shared__ float sm[];                     // [k*k weights][(TILE + k - 1)^2 input tile]
  const int r = k >> 1, tw = DW_TILE + k - 1;
  float* sw = sm;
  float* st = sm + k * k;
  const float* plane = x + (int64_t)blockIdx.z * h * w;
  const int x0 = blockIdx.x * DW_TILE, y0 = blockIdx.y * DW_TILE;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) sw[i] = wt[i];
  for (int i = threadIdx.x; i < tw * tw; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int iy = y0 + ty - r, ix = x0 + tx - r;
    st[i] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(plane + (int64_t)iy * w + ix) : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
  for (int ly = ly0; ly < DW_TILE; ly += 8) {
    const int oy = y0 + ly, ox = x0 + lx;
    if (oy >= h || ox >= w) continue;
    float acc = 0.f;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) acc = fmaf(st[(ly + i) * tw + lx + j], sw[i * k + j], acc);
    if (min_with_input) acc = fminf(acc, st[(ly + r) * tw + lx + r]);
    out[(int64_t)blockIdx.z * h * w + (int64_t)oy * w + ox] = acc;
  }
}

// max over the elements below the threshold (float max through the ordered-int trick; values may be negative)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void __launch_bounds__(256) below_threshold_max_kernel(const float* __restrict__ x, float thr, float* __restrict__ mx, int64_t n) {
  float m = -INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    if (!(v >= thr)) m = fmaxf(m, v);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_float(mx, m);
}
__global__ void __launch_bounds__(256) soft_erosion_finish_kernel(float* __restrict__ x, uint8_t* __restrict__ mask, float thr, const float* __restrict__ mx,
                                                                  int64_t n) {
  const float m = __ldg(mx);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const bool in = v >= thr;
    x[i] = in ? 1.f : v / m;
    mask[i] = in ? 1 : 0;
  }
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// cv2.pyrDown: out[y,x] = sum_{i,j} k[i]k[j] in[2y+i-2, 2x+j-2] / 256, k = [1 4 6 4 1], BORDER_REFLECT_101; round_u8: floor((sum+128)/256)
__global__ void __launch_bounds__(256) pyr_down_kernel(const float* __restrict__ x, float* __restrict__ out, int h, int w, int oh, int ow, int round_u8,
                                                       int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const int64_t t = i / ow;
    const int oy = (int)(t % oh);
    const float* plane = x + (t / oh) * (int64_t)h * w;
    int xs[5], ys[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      xs[d] = reflect101(2 * ox + d - 2, w);
      ys[d] = reflect101(2 * oy + d - 2, h);
    }
    float row[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const float* rp = plane + (int64_t)ys[d] * w;
      row[d] = __ldg(rp + xs[2]) * 6.f + (__ldg(rp + xs[1]) + __ldg(rp + xs[3])) * 4.f + __ldg(rp + xs[0]) + __ldg(rp + xs[4]);
    }
    const float s = row[2] * 6.f + (row[1] + row[3]) * 4.f + row[0] + row[4];
    out[i] = round_u8 ? floorf((s + 128.f) * (1.f / 256.f)) : s * (1.f / 256.f);
  }
}

// cv2.pyrUp: zero-insert x2 and [1 4 6 4 1]/8 per axis; first sample's left / top neighbour reflected (101), last sample's right / bottom
// neighbour replicated.  mode 0: out = up; 1: out = other - up (Laplacian level); 2: out = up + other (reconstruction)
__device__ __forceinline__ float pyr_up_row(const float* __restrict__ rp, int sx, int px, int w) {
  const float c = __ldg(rp + sx);
  const float r = __ldg(rp + (sx + 1 < w ? sx + 1 : sx));
  if (px) return (c + r) * 4.f;
  const float l = __ldg(rp + (sx > 0 ? sx - 1 : (w > 1 ? 1 : 0)));
  return l + c * 6.f + r;
}
__global__ void __launch_bounds__(256) pyr_up_kernel(const float* __restrict__ x, const float* __restrict__ other, float* __restrict__ out, int h, int w,
                                                     int mode, int64_t total) {
  const int ow = 2 * w, oh = 2 * h;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const int64_t t = i / ow;
    const int oy = (int)(t % oh);
    const float* plane = x + (t / oh) * (int64_t)h * w;
    const int sx = ox >> 1, px = ox & 1, sy = oy >> 1, py = oy & 1;
    const float c = pyr_up_row(plane + (int64_t)sy * w, sx, px, w);
    const float d = pyr_up_row(plane + (int64_t)(sy + 1 < h ? sy + 1 : sy) * w, sx, px, w);
    float v;
    if (py) {
      v = (c + d) * 4.f;
    } else {
      const float u = pyr_up_row(plane + (int64_t)(sy > 0 ? sy - 1 : (h > 1 ? 1 : 0)) * w, sx, px, w);
      v = u + c * 6.f + d;
    }
    v *= (1.f / 64.f);
    if (mode == 1) v = __ldg(other + i) - v;
    else if (mode == 2) v = v + __ldg(other + i);
    out[i] = v;
  }
}

// ls = la * gm + lb * (1 - gm); gm has the images' shape, or one plane per sample (m_planes_per_sample = 1) broadcast over channels
__global__ void __launch_bounds__(256) pyr_blend_kernel(const float* __restrict__ la, const float* __restrict__ lb, const float* __restrict__ gm,
                                                        float* __restrict__ out, int64_t hw, int channels, int m_channels, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t mi = i;
    if (m_channels != channels) {
      const int64_t plane = i / hw;
      mi = (plane / channels) * hw + (i - plane * hw);
    }
    const float g = __ldg(gm + mi);
    out[i] = __ldg(la + i) * g + __ldg(lb + i) * (1.0f - g);
  }
}

}  // namespace e4s

using namespace e4s;

extern "C" int e4s_depthwise_conv_f32(const float* x, const float* weight, float* out, int64_t planes, int h, int w, int k, int min_with_input,
                                      void* stream) {
  E4S_REQUIRE(x && weight && out && planes > 0 && planes < 65536 && h > 0 && w > 0 && k > 0 && (k & 1) && k <= 31, "depthwise_conv: bad args (odd k <= 31)");
  const int tw = DW_TILE + k - 1;
  const size_t smem = (size_t)(k * k + tw * tw) * sizeof(float);
  dim3 grid(ceil_div(w, DW_TILE), ceil_div(h, DW_TILE), (unsigned)planes);
  depthwise_conv_kernel<<<grid, DW_TILE * 8, smem, as_stream(stream)>>>(x, weight, out, h, w, k, min_with_input);
  return check_launch("depthwise_conv");
}

extern "C" int e4s_soft_erosion_finish_f32(float* x, uint8_t* mask, int64_t n, float threshold, float* scratch_max, void* stream) {
  E4S_REQUIRE(x && mask && scratch_max && n > 0, "soft_erosion_finish: bad args");
  cudaStream_t s = as_stream(stream);
  const float ninf = -INFINITY;
  cudaError_t e = cudaMemcpyAsync(scratch_max, &ninf, sizeof(float), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "soft_erosion_finish: %s", cudaGetErrorString(e));
  below_threshold_max_kernel<<<pb_grid(n), 256, 0, s>>>(x, threshold, scratch_max, n);
  int rc = check_launch("soft_erosion_max");
  if (rc) return rc;
  soft_erosion_finish_kernel<<<pb_grid(n), 256, 0, s>>>(x, mask, threshold, scratch_max, n);
  return check_launch("soft_erosion_finish");
}

extern "C" int e4s_pyr_down_f32(const float* x, float* out, int64_t planes, int h, int w, int round_u8, void* stream) {
  E4S_REQUIRE(x && out && planes > 0 && h > 0 && w > 0, "pyr_down: bad args");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const int64_t total = planes * oh * ow;
  pyr_down_kernel<<<pb_grid(total), 256, 0, as_stream(stream)>>>(x, out, h, w, oh, ow, round_u8, total);
  return check_launch("pyr_down");
}

extern "C" int e4s_pyr_up_f32(const float* x, const float* other, float* out, int64_t planes, int h, int w, int mode, void* stream) {
  E4S_REQUIRE(x && out && planes > 0 && h > 0 && w > 0 && mode >= 0 && mode <= 2 && (mode == 0 || other), "pyr_up: bad args");
  const int64_t total = planes * 4 * (int64_t)h * w;
  pyr_up_kernel<<<pb_grid(total), 256, 0, as_stream(stream)>>>(x, other, out, h, w, mode, total);
  return check_launch("pyr_up");
}

extern "C" int e4s_pyr_blend_f32(const float* la, const float* lb, const float* gm, float* out, int batch, int channels, int m_channels, int h, int w,
                                 void* stream) {
  E4S_REQUIRE(la && lb && gm && out && batch > 0 && channels > 0 && h > 0 && w > 0 && (m_channels == channels || m_channels == 1), "pyr_blend: bad args");
  const int64_t hw = (int64_t)h * w, total = (int64_t)batch * channels * hw;
  pyr_blend_kernel<<<pb_grid(total), 256, 0, as_stream(stream)>>>(la, lb, gm, out, hw, channels, m_channels, total);
  return check_launch("pyr_blend");
}

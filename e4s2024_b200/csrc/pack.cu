// Weight re-layout kernels: nn.Module parameter tensors -> the engine's [phase][K = tap*cin + ci][cout_pad] fp32 matrices, one launch per
// layer (the first version did this with torch elementwise ops: ~2300 tiny launches for the eight up-convolutions of a generator).
#include "common.cuh"

namespace e4s {

// w [cout][cin][kh][kw] -> out [kh*kw*cin_pad][cout_pad], out[(tap*cin_pad + ci)][co] = scale * w[co][ci][tap]; zero in the padding.
// sumsq != 0: out [cin_pad][cout_pad] = sum over taps of (scale * w)^2   (the demodulation table GEMM's weight, model.py:279-281)
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps,
                                                                int cin_pad, int cout_pad, float scale, int sumsq, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    const int64_t k = i / cout_pad;
    float v = 0.f;
    if (sumsq) {
      const int ci = (int)k;
      if (co < cout && ci < cin) {
        const float* p = w + ((int64_t)co * cin + ci) * taps;
        for (int t = 0; t < taps; ++t) {
          const float s = scale * p[t];
          v = fmaf(s, s, v);
        }
      }
    } else {
      const int ci = (int)(k % cin_pad), tap = (int)(k / cin_pad);
      if (co < cout && ci < cin) v = scale * w[((int64_t)co * cin + ci) * taps + tap];
    }
    out[i] = v;
  }
}

// conv_transpose2d(stride 2, weight w[co][ci][3][3] used as [ci][co][3][3]) followed by upfirdn2d(fir 4x4, pad=(1,1)) == four 3x3 phase
// filters at input resolution:  out[2A+py, 2B+px] = sum_{u,v} x[A-1+u, B-1+v] * Wph[py,px,u,v],
//   Wph[py,px,u,v] = sum_{m,n} w[ky,kx] * fir[3-m][3-n],  ky = 2(1-u) + py + m - 1, kx = 2(1-v) + px + n - 1, 0 <= ky,kx <= 2.
// out [4 = py*2+px][(u*3+v)*cin + ci][cout_pad].
__global__ void __launch_bounds__(256) pack_upconv_weights_kernel(const float* __restrict__ w, const float* __restrict__ fir, float* __restrict__ out,
                                                                  int cout, int cin, int cout_pad, float scale, int64_t total) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = fir[15 - threadIdx.x];          // flipped in both axes
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    int64_t t = i / cout_pad;
    const int ci = (int)(t % cin);
    t /= cin;
    const int v = (int)(t % 3), u = (int)((t / 3) % 3), ph = (int)(t / 9);
    const int py = ph >> 1, px = ph & 1;
    float acc = 0.f;
    if (co < cout) {
      const float* wp = w + ((int64_t)co * cin + ci) * 9;
      for (int m = 0; m < 4; ++m) {
        const int ky = 2 * (1 - u) + py + m - 1;
        if (ky < 0 || ky > 2) continue;
        for (int n = 0; n < 4; ++n) {
          const int kx = 2 * (1 - v) + px + n - 1;
          if (kx < 0 || kx > 2) continue;
          acc += (scale * wp[ky * 3 + kx]) * kf[m * 4 + n];
        }
      }
    }
    out[i] = acc;
  }
}

static inline unsigned pack_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  return (unsigned)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

}  // namespace e4s

using namespace e4s;

extern "C" int e4s_pack_conv_weights_f32(const float* w, float* out, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, float scale,
                                         int sumsq, void* stream) {
  E4S_REQUIRE(w && out && cout > 0 && cin > 0 && kh > 0 && kw > 0 && cin_pad >= cin && cout_pad >= cout, "pack_conv_weights: bad args");
  const int64_t total = (int64_t)(sumsq ? 1 : kh * kw) * cin_pad * cout_pad;
  pack_conv_weights_kernel<<<pack_grid(total), 256, 0, as_stream(stream)>>>(w, out, cout, cin, kh * kw, cin_pad, cout_pad, scale, sumsq, total);
  return check_launch("pack_conv_weights");
}

extern "C" int e4s_pack_upconv_weights_f32(const float* w, const float* fir, float* out, int cout, int cin, int cout_pad, float scale, void* stream) {
  E4S_REQUIRE(w && fir && out && cout > 0 && cin > 0 && cout_pad >= cout, "pack_upconv_weights: bad args");
  const int64_t total = (int64_t)4 * 9 * cin * cout_pad;
  pack_upconv_weights_kernel<<<pack_grid(total), 256, 0, as_stream(stream)>>>(w, fir, out, cout, cin, cout_pad, scale, total);
  return check_launch("pack_upconv_weights");
}

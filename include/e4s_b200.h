/* e4s_b200.h -- C-ABI of the B200-native E4S hot path (libe4s_b200.so).
 *
 * Boundary contract:
 *  - every function is extern "C", takes raw DEVICE pointers + int shapes + a cudaStream_t
 *    (passed as void*), returns 0 on success or a negative E4S_ERR_* code; the message is
 *    available from e4s_last_error() (thread-local).
 *  - the caller owns every buffer (inputs, outputs, packed weights, workspaces); the library
 *    never allocates, frees or retains caller memory.  All work is enqueued on `stream`
 *    (CUDA-graph capturable); no host synchronisation inside.
 *  - there is no CPU fallback: without a CUDA device every entry point fails.
 *
 * Which reference interface each entry point replaces (paths under the reference root):
 *  e4s_upfirdn2d_f32      <- pybind `upfirdn2d(input,kernel,up_x,up_y,down_x,down_y,pad_x0,pad_x1,pad_y0,pad_y1)`
 *                            models/stylegan2/op/upfirdn2d.cpp:12-23, kernel upfirdn2d_kernel.cu:52-137
 *  e4s_bias_act_f32       <- pybind `fused_bias_act(input,bias,refer,act,grad,alpha,scale)`
 *                            models/stylegan2/op/fused_bias_act.cpp:11-21, kernel fused_bias_act_kernel.cu:18-49
 *  e4s_conv_f32 / e4s_conv_tc  <- torch F.conv2d / F.conv_transpose2d call sites of the path:
 *                            models/stylegan2/op/conv2d_gradfix.py:34-42,66-75 (ModulatedConv2d, model.py:287-318),
 *                            models/encoders/helpers.py:128-139, psp_encoders.py:334,
 *                            swap_face_fine/face_parsing/model.py:23-35, resnet.py:17-49,62-64, F.linear model.py:156-162
 *  e4s_torgb_f32          <- ToRGB.forward, models/stylegan2/model.py:439-479 (1x1 modconv + bias + Upsample(skip))
 *  e4s_mask_labels        <- F.interpolate(mask,'nearest') region lookup, model.py:391,445; psp_encoders.py:357
 *  e4s_chan_stats_f32     <- InstanceNorm2d statistics / AdaptiveAvgPool2d / F.avg_pool2d(full), helpers.py:60,128-138
 *  e4s_vec_fc_f32         <- SEModule fc1/fc2 (helpers.py:61-71), conv_atten/conv_avg/FFM 1x1 on pooled vectors (model.py:86-88,117-118,208-213)
 *  e4s_residual_combine_f32 <- bottleneck_IR_SE_Ours tail `res*gate + shortcut` (helpers.py:141-144), ARM/FFM gating (model.py:89,130,214-215)
 *  e4s_masked_mean_f32    <- FSEncoder_PSP.get_per_comp_styleCode, psp_encoders.py:355-375
 *  e4s_resize_bilinear_f32<- F.interpolate(img,(256,256),'bilinear') networks.py:114,217; BiSeNet head upsample model.py:257-259
 *  e4s_maxpool3x3s2_f32   <- nn.MaxPool2d(3,2,1), resnet.py:66,75
 *  e4s_upsample_argmax_u8 <- F.interpolate(...,align_corners=True) + torch.argmax + seg19->seg12 LUT,
 *                            model.py:257, face_parsing_demo.py:170, datasets/dataset.py:58-108
 *  e4s_bicubic_down_norm_f32 <- BicubicDownSample.forward + clamp + (x-mean)/std, face_parsing_demo.py:46-84,155
 *  e4s_nchw_to_nhwc_f32 / e4s_nhwc_to_nchw_f32 <- layout adapters at the nn.Module boundary (API is NCHW)
 */
#ifndef E4S_B200_H
#define E4S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E4S_OK 0
#define E4S_ERR_ARG (-1)      /* bad shape / alignment / null pointer */
#define E4S_ERR_CUDA (-2)     /* CUDA runtime error (launch failure, no device) */
#define E4S_ERR_UNSUPPORTED (-3)

/* activation codes for E4SConv.act */
enum { E4S_ACT_NONE = 0, E4S_ACT_LRELU = 1, E4S_ACT_RELU = 2, E4S_ACT_PRELU = 3, E4S_ACT_SIGMOID = 4,
       E4S_ACT_RSQRT_EPS = 5 };
/* E4SConv.tc_fmt */
enum { E4S_TC_BF16 = 0, E4S_TC_F16 = 1 };
/* E4SConv.mode */
enum { E4S_CONV_NORMAL = 0, E4S_CONV_UP2_POLYPHASE = 1 };

/* One convolution / GEMM launch.  Activations are fp32 NHWC with an arbitrary pixel pitch
 * (so channel slices of wider buffers can be read / written).  GEMM view: M = B*Hout*Wout
 * output pixels, N = cout, K = taps*cin with k = tap*cin + ci.
 *
 * A operand (gathered on the fly, never materialised):
 *   a[p, tap, ci] = T( x[b, iy, ix, ci] ),  iy = oy*stride - pad + ky  (then >> in_shift: the conv
 *   reads a nearest-upsampled view of x),  zero outside the image.
 *   T applies, in order: InstanceNorm (v - in_mean[b,ci]) * in_rstd[b,ci]   (if in_mean)
 *                        square v*v (if in_square; used for the demodulation table GEMM)
 *                        style modulation v * smod[(b*regions + r(p))*cin + ci] (if smod)
 *   r(p) = labels[b, floor(oy*lab_h/Hout), floor(ox*lab_w/Wout)]   (0 if labels == NULL)
 * mode E4S_CONV_UP2_POLYPHASE: conv_transpose2d(stride 2) + 4x4 FIR folded into four 3x3 phase
 *   filters; out pixel (2a+py, 2b+px) gathers x[a-1+u, b-1+v]; weights hold 4 phase matrices.
 *
 * Epilogue, in order:  v = acc * demod[(b*regions + r)*cout + n]      (if demod)
 *                      v *= pixw[b*pixw_sb + sy*lab_w + sx]             (if pixw; generic float masks)
 *                      v  = v * ch_scale[n]                             (if ch_scale)
 *                      v += noise_w[0] * noise[b*noise_sb + n*noise_sc + oy*Wout + ox] (if noise)
 *                      v += ch_shift[n]                                 (if ch_shift)
 *                      v += res[pix*res_pitch + n]                      (if res && !res_after_act)
 *                      v  = act(v)   LRELU: (v<0 ? v*act_slope : v)*act_gain; PRELU: slope = act_prelu[n];
 *                                     RSQRT_EPS: rsqrt(v + act_slope)
 *                      v += res[...]                                    (if res && res_after_act)
 *                      out[pix*out_pitch + n] = v   (or += if accumulate)
 */
typedef struct E4SConv {
  const float* x;   int64_t x_pitch;       /* floats per pixel of the input buffer */
  int32_t batch, hin, win, cin;            /* cin % 8 == 0 */
  int32_t in_shift;                        /* log2 nearest-upsample factor of the input view */
  int32_t in_square;                       /* square the gathered value (after InstanceNorm) */
  const float* w;                          /* packed weights, see e4s_conv_f32 / e4s_conv_tc */
  int32_t cout, cout_pad;                  /* cout_pad % 4 == 0, row length of packed w */
  int32_t kh, kw, stride, pad, mode;
  int32_t hout, wout;
  const float* in_mean; const float* in_rstd;          /* [batch, cin] or NULL */
  const float* smod; const uint8_t* labels;            /* [batch, regions, cin]; [batch, lab_h, lab_w] */
  int32_t regions, lab_h, lab_w;
  const float* demod;                                  /* [batch, regions, cout] or NULL */
  const float* pixw; int64_t pixw_sb;                  /* float mask plane(s), sampled like labels */
  const float* ch_scale; const float* ch_shift;        /* [cout] or NULL */
  const float* noise; const float* noise_w; int64_t noise_sb, noise_sc;
  const float* res; int64_t res_pitch; int32_t res_after_act;
  int32_t act; float act_slope, act_gain; const float* act_prelu;
  float* out; int64_t out_pitch; int32_t accumulate;
  /* Fused ToRGB tail (reference models/stylegan2/model.py:439-479 applied to this layer's activated output without
   * re-reading it; e4s_conv_tc only, un-masked same-resolution 3x3 layers with cout <= 128):
   *   rgb[b,c,y,x] = sum_co act_out[b,y,x,co] * rgb_w[c,co] * rgb_smod[b,co] + rgb_bias[c] + FIR-upsample(rgb_skip)[b,c,y,x]
   * rgb / rgb_skip are NCHW fp32 ([batch,3,hout,wout] / [batch,3,hout/2,wout/2]); rgb_fir is the 4x4 kernel of
   * Upsample (upfirdn2d up=2, pad=(2,1)).  With rgb set, out may be NULL: the activations are then never written. */
  float* rgb; const float* rgb_w; const float* rgb_smod; const float* rgb_bias; const float* rgb_skip; const float* rgb_fir;
  /* Device-side launch predicate (NULL = always run): the kernel does its work only if (*pred_count > pred_limit) == (pred_run_if_gt != 0),
   * otherwise every CTA exits at once.  Lets the host enqueue BOTH candidate kernels of a masked layer (per-(tile, region) jobs on the
   * halo kernel vs. the per-row gather / wide kernel) without reading the job count back: no host synchronisation inside a forward. */
  const int32_t* pred_count; int32_t pred_limit; int32_t pred_run_if_gt;
  /* e4s_conv_tc only.  tc_fmt: operand format of the 3-pass split (must match the packed weights): E4S_TC_BF16 = two bf16 per value
   * (2^-17 relative, any fp32 range: the StyleGAN2 / encoder layers), E4S_TC_F16 = two fp16 per value (2^-22 relative = fp32 class,
   * activations must stay below 65504: BiSeNet, whose label map has to match the reference's argmax).  tc_out_scale multiplies the
   * accumulator (0 = 1: undoes the power-of-two pre-scale e4s_pack_weights_tc_fmt applied to fp16 weights).  tc_unbias: relative
   * correction per tcgen05.mma accumulate step for the truncating fp32 accumulate of the tensor core (0 = the library's measured
   * default for the format, < 0 = none; tests/micro/acc_bias.py measures it). */
  int32_t tc_fmt; float tc_out_scale; float tc_unbias;
} E4SConv;

const char* e4s_last_error(void);
/* number of kernels this library has launched in this process (bench.py `gpu_launches`) */
int64_t e4s_launch_count(void);
int e4s_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* sizeof(struct E4SConv) as compiled, so bindings can verify their mirror of the struct */
int e4s_sizeof_conv(void);

/* fp32 CUDA-core implicit GEMM.  w layout: [phase][K][cout_pad] (phase = 1 or 4). */
int e4s_conv_f32(const E4SConv* p, void* stream);
/* `count` independent E4S_CONV_NORMAL problems (host array) in one launch per 32 problems: blockIdx.z = problem */
int e4s_conv_f32_batched(const E4SConv* params, int count, void* stream);

/* tcgen05 (5th-gen tensor core) implicit GEMM with the 3-pass bf16 hi/lo split (fp32 accumulate in TMEM).
 * w points to the packed bf16 image produced by e4s_pack_weights_tc; cin % 8 == 0, cout in {32,64,128,256*n}. */
int e4s_conv_tc(const E4SConv* p, const void* w_packed, void* stream);
/* Split-K form of e4s_conv_tc for layers with too few output tiles to fill the machine (the generator's 4^2 - 16^2 layers: K = 4608 over
 * batch*16 ... batch*256 output pixels): ksplit CTAs per 128-pixel tile each accumulate 1/ksplit of the 64-wide K chunks into
 * ws[z][pixel][cout] (raw fp32 accumulators), a second kernel adds the slices in index order (deterministic) and applies the epilogue.
 * Restrictions: mode E4S_CONV_NORMAL, bf16 split, the piecewise-linear epilogue family (no residual / float mask / accumulate / per-channel
 * noise), no fused ToRGB, ksplit in [2, 64] dividing ceil(kh*kw*cin / 64).  ws: e4s_conv_tc_splitk_ws_bytes bytes, 16-byte aligned. */
int64_t e4s_conv_tc_splitk_ws_bytes(const E4SConv* p, int ksplit);
int e4s_conv_tc_splitk(const E4SConv* p, const void* w_packed, int ksplit, void* ws, void* stream);
/* Regional (masked) 3x3 / up-conv layer through the halo kernel: `jobs` holds one int4 (b, y0, x0, region) per (16x8 tile,
 * region present in the tile), built on the device by e4s_region_tile_jobs; job_count is the device counter, job_count_host
 * its value.  Every output pixel is written by exactly one job (the one of its own region). */
int e4s_conv_tc_regions(const E4SConv* p, const void* w_packed, const int32_t* jobs, const int32_t* job_count, int job_count_host,
                        void* stream);
/* labels u8 [batch, lab_h, lab_w] -> job list for an hout x wout layer (up2: tiles over the hout/2 x wout/2 input grid);
 * *count must be zeroed by the caller and receives the number of jobs (entries beyond max_jobs are dropped) */
int e4s_region_tile_jobs(const uint8_t* labels, int batch, int lab_h, int lab_w, int hout, int wout, int up2, int32_t* jobs,
                         int32_t* count, int max_jobs, void* stream);
/* ---- Regional up-convolution at its algorithmic cost (csrc/conv_tc_upz.cu; reference models/stylegan2/model.py:287-300,395-398) ------
 * conv_transpose2d(stride 2) runs as a GEMM over (cell, region) rows -- cell (cy,cx), 0 <= cy <= hin, 0 <= cx <= win, holds the four
 * values z[2cy+py, 2cx+px] of the (2hin+1) x (2win+1) transposed-convolution output and reads the 2x2 input window x[cy-1..cy, cx-1..cx];
 * one row per region whose output pixels' 4x4 blur windows touch the cell -- and the 4x4 FIR + noise + bias + activation as a finishing
 * pass: 9 instead of 36 cin*cout MACs per input pixel on masks with few regions per cell.
 * e4s_upz_build_rows: labels u8 [batch, lab_h, lab_w] -> cells int32 [batch, hin+1, win+1, 2] = (region bit mask, first row) and rows
 *   int32 [max_rows, 2] = (b << 8 | region, cy << 16 | cx); *count (zeroed by the caller) receives the number of rows, entries beyond
 *   max_rows are dropped (both kernels of e4s_conv_tc_upz then do nothing: the caller runs e4s_conv_tc instead, E4SConv.pred_*).
 * e4s_pack_convt_weights_f32: w [cout][cin][3][3] -> out [9][cin][cout_pad] (the 9 taps as [cin x cout] matrices in the order the kernel
 *   consumes them), to be packed with e4s_pack_weights_tc(out, phases = 9, k = cin, cin, cout, cout_pad).
 * e4s_conv_tc_upz: p as for e4s_conv_tc with mode E4S_CONV_UP2_POLYPHASE, labels, smod (cin % 64 == 0, cout % 128 == 0, bf16 split,
 *   single-channel noise, NONE / LRELU); fir = the 4x4 blur kernel (Blur(pad=(1,1)), already x4); z: scratch of max_rows*4*cout floats.
 *   Un-masked layers (labels == NULL, regions == 1; cout in {32, 64, 128 n}): DIRECT form, cells = rows = count = NULL and
 *   max_rows = batch*(hin+1)*(win+1) -- row i is cell i, no list is built. */
int e4s_upz_build_rows(const uint8_t* labels, int batch, int lab_h, int lab_w, int hin, int win, int32_t* cells, int32_t* rows,
                       int32_t* count, int max_rows, void* stream);
int e4s_pack_convt_weights_f32(const float* w, float* out, int cout, int cin, int cout_pad, float scale, void* stream);
int e4s_conv_tc_upz(const E4SConv* p, const void* w_packed9, const float* fir, const int32_t* cells, const int32_t* rows,
                    const int32_t* count, int max_rows, float* z, void* stream);
/* bytes needed for the packed tensor-core weights of a [phases, K, cout] fp32 matrix */
int64_t e4s_pack_weights_tc_bytes(int phases, int k, int cout);
/* w_f32: [phases][K = taps*cin][cout_pad] (the e4s_conv_f32 layout) -> w_packed (hi/lo bf16, UMMA K-major SW128 tiles;
 * 64-wide K chunks ordered channel-group-major when cin % 64 == 0) */
/* debug aid: record a (role, job, event, clock64) timeline of CTA 0 of the halo kernel into buf (2 x u64 per record); NULL removes it */
int e4s_debug_halo_trace(void* buf, int cap_records);
/* debug aid: profiling experiments on the halo kernel (bit0 skip epilogue math+stores, bit1 skip tcgen05.ld, bit2 skip MMAs); 0 = normal */
int e4s_debug_halo_flags(int flags);
/* debug aid: profiling experiments on conv_tc_upz_kernel (bit0 skip the Z stores, bit1 skip the MMAs, bit2 skip the A gather loads, bit3 skip the TMEM reads); 0 = normal */
int e4s_debug_upz_flags(int flags);
int e4s_pack_weights_tc(const float* w_f32, int phases, int k, int cin, int cout, int cout_pad, void* w_packed, void* stream);
/* same with an explicit operand format (E4S_TC_*): every weight is multiplied by `scale` (a power of two chosen by the caller so that
 * max|w|*scale ~ 2^14: keeps the fp16 lo parts out of the subnormal range) before the hi/lo split; the launch passes 1/scale as
 * E4SConv.tc_out_scale.  fmt = E4S_TC_BF16, scale = 1 is e4s_pack_weights_tc. */
int e4s_pack_weights_tc_fmt(const float* w_f32, int phases, int k, int cin, int cout, int cout_pad, int fmt, float scale, void* w_packed,
                            void* stream);

/* nn.Module weights -> the engine's fp32 matrices, one launch per layer.
 * e4s_pack_conv_weights_f32: w [cout][cin][kh][kw] (nn.Conv2d / F.linear weight with kh = kw = 1) -> out [kh*kw*cin_pad][cout_pad],
 *   out[tap*cin_pad + ci][co] = scale * w[co][ci][tap], zero in the channel padding.  sumsq != 0: out [cin_pad][cout_pad] =
 *   sum_taps (scale*w)^2, the weight of the demodulation table GEMM (reference models/stylegan2/model.py:279-281).
 * e4s_pack_upconv_weights_f32: w [cout][cin][3][3] + the 4x4 blur kernel -> the four 3x3 phase filters of conv_transpose2d(stride 2)
 *   + Blur(pad=(1,1)) (model.py:287-300), out [4][9*cin][cout_pad]. */
int e4s_pack_conv_weights_f32(const float* w, float* out, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, float scale,
                              int sumsq, void* stream);
int e4s_pack_upconv_weights_f32(const float* w, const float* fir, float* out, int cout, int cin, int cout_pad, float scale, void* stream);

/* upfirdn2d on NCHW fp32 [planes, in_h, in_w] (planes = B*C). kernel [kh,kw] is correlated FLIPPED. */
int e4s_upfirdn2d_f32(const float* x, const float* kernel, float* out, int64_t planes, int in_h, int in_w,
                      int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                      int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream);

/* y[i] = lrelu(x[i] + bias[(i / inner) % channels], slope) * scale   (bias may be NULL) */
int e4s_bias_act_f32(const float* x, const float* bias, float* y, int64_t n, int64_t inner, int channels,
                     float slope, float scale, void* stream);

/* gradient form of e4s_bias_act_f32 (reference fused_bias_act_kernel.cu act=3, grad=1): y[i] = (ref[i] > 0 ? v : v*slope) * scale with
 * v = g[i] + bias[channel] (bias may be NULL); ref = the forward output.  Used by the autograd of fused_leaky_relu (op/fused_act.py:18-69). */
int e4s_bias_act_grad_f32(const float* g, const float* bias, const float* ref, float* y, int64_t n, int64_t inner, int channels,
                          float slope, float scale, void* stream);

/* NHWC noise+bias+leaky-relu for the generic-mask path: y = lrelu(x + nw*noise + bias[c]) * scale, in place ok */
int e4s_noise_bias_act_nhwc_f32(float* x, int batch, int h, int w, int c, const float* noise, const float* noise_w,
                                int64_t noise_sb, int64_t noise_sc, const float* bias, float slope, float scale,
                                void* stream);

int e4s_nchw_to_nhwc_f32(const float* x, float* y, int batch, int c, int h, int w, int c_pad, void* stream);
int e4s_nhwc_to_nchw_f32(const float* x, int64_t x_pitch, float* y, int batch, int c, int h, int w, void* stream);

/* mask [batch, k, h, w] float -> labels u8 [batch, h, w]: index of the single non-zero entry (which must
 * equal 1).  flags[0] counts thread blocks that saw a pixel NOT of that form (soft, overlapping or empty
 * masks) -> the caller must take the generic per-region path. */
int e4s_mask_labels(const float* mask, int batch, int k, int h, int w, uint8_t* labels, int32_t* flags, void* stream);

/* ToRGB: rgb[b,c,y,x] = sum_ci x[b,y,x,ci] * smod[(b*regions+r)*cin + ci] * wrgb[c*cin + ci] + bias[c] + up2(skip)[b,c,y,x]
 * (wrgb = scale * weight[3,cin]; skip [batch,3,h/2,w/2] NCHW or NULL, 4x4 FIR `fir` as in Upsample: up=2, pad=(2,1)).
 * pixw != NULL: multiply the dot product by the float mask (generic path); accumulate != 0 adds into rgb
 * without bias/skip. */
int e4s_torgb_f32(const float* x, int64_t x_pitch, int batch, int h, int w, int cin, const float* smod, const float* wrgb,
                  const uint8_t* labels, int regions, int lab_h, int lab_w, const float* pixw, int64_t pixw_sb,
                  const float* bias, const float* skip, const float* fir, float* rgb, int accumulate, void* stream);

/* per-(b,c) statistics of an NHWC tensor: mean, rstd = 1/sqrt(var_biased + eps) (either may be NULL).
 * ws: workspace of e4s_chan_stats_ws_bytes(batch, c) bytes whose first 64 KB (arrival counters of the single-launch form) must be ZERO
 * when the call starts; the call leaves them zero again, so one zero-filled buffer serves any number of calls on a stream.
 * Deterministic (fixed split, fp64 partials, slices reduced in index order). */
int64_t e4s_chan_stats_ws_bytes(int batch, int c);
int e4s_chan_stats_f32(const float* x, int64_t x_pitch, int batch, int hw, int c, float eps, float* mean, float* rstd,
                       void* ws, void* stream);

/* y[b,o] = act( scale[o] * sum_i w[o*cin + i] * x[b*x_stride + i] + shift[o] )  -- tiny per-sample FC */
int e4s_vec_fc_f32(const float* x, int64_t x_stride, const float* w, const float* scale, const float* shift,
                   float* y, int batch, int cin, int cout, int act, void* stream);

/* out[b,p,c] = T(a[b,p,c]) * gate[b,c] (+ T2(r[b, p', c]))   with
 *   T  = (v - a_mean[b,c]) * a_rstd[b,c] if a_mean else v ;  gate == NULL -> 1 ; gate_plus_one -> v*gate + v
 *   r [batch, r_h, r_w, r_pitch]: r_sub > 1 reads pixel (y*r_sub, x*r_sub) (MaxPool2d(1, stride) shortcut);
 *   r_sub == 1 reads the legacy-nearest resize of r to h x w (identity when sizes match).
 *   T2 normalises with r_mean/r_rstd if given.  Then relu (if set), then PReLU with slopes prelu[c] (if given). */
int e4s_residual_combine_f32(const float* a, int64_t a_pitch, const float* a_mean, const float* a_rstd,
                             const float* gate, int gate_plus_one, const float* r, int64_t r_pitch, int r_h, int r_w,
                             int r_sub, const float* r_mean, const float* r_rstd, int relu, const float* prelu,
                             float* out, int64_t out_pitch, int batch, int h, int w, int c, void* stream);

/* codes[b, k, c_off + c] = mean over pixels with mask[b,k,sy,sx] != 0 of feat[b,y,x,c]; 0 if none */
int e4s_masked_mean_f32(const float* feat, int64_t f_pitch, int batch, int h, int w, int c, const float* mask, int k,
                        int mh, int mw, float* codes, int64_t codes_stride_b, int64_t codes_stride_k, int c_off,
                        void* stream);

/* The same pooling for several feature maps of one forward: the mask is reduced once to a membership map, bits[b,y,x] bit j set <=>
 * mask[b,j,y,x] != 0 (k <= 32), and each call reads one word per feature pixel.  ws: e4s_masked_mean_ws_bytes(batch, c, k) bytes
 * (fp64 partial sums of 8 pixel chunks, reduced in a fixed order: batch-invariant results). */
int e4s_mask_member_bits_u32(const float* mask, int batch, int k, int h, int w, uint32_t* bits, void* stream);
int64_t e4s_masked_mean_ws_bytes(int batch, int c, int k);
int e4s_masked_mean_bits_f32(const float* feat, int64_t f_pitch, int batch, int h, int w, int c, const uint32_t* bits, int k, int mh,
                             int mw, float* codes, int64_t codes_stride_b, int64_t codes_stride_k, int c_off, void* ws, void* stream);

/* bilinear resize NCHW -> NHWC (c_pad channels, zero filled) or NHWC->NCHW; align_corners as torch */
int e4s_resize_bilinear_nchw_to_nhwc_f32(const float* x, int batch, int c, int hin, int win, float* y, int hout,
                                         int wout, int c_pad, int align_corners, void* stream);
int e4s_resize_bilinear_nhwc_to_nchw_f32(const float* x, int64_t x_pitch, int batch, int c, int hin, int win, float* y,
                                         int hout, int wout, int align_corners, void* stream);
int e4s_maxpool3x3s2_nhwc_f32(const float* x, int batch, int h, int w, int c, float* y, void* stream);

/* logits NHWC [batch, hin, win, pitch] (first c channels) -> bilinear(align_corners=True) upsample to
 * hout x wout, argmax over c (lowest index wins ties), optional 256-entry LUT; labels u8 [batch,hout,wout] */
int e4s_upsample_argmax_u8(const float* logits, int64_t pitch, int batch, int c, int hin, int win, int hout, int wout,
                           const uint8_t* lut, uint8_t* labels, void* stream);

/* FaceParser front-end: x NCHW [batch,3,hin,win] in [0,1] -> separable `factor*4`-tap bicubic down by `factor`
 * (reflect pad), clamp(0,1) if do_clamp, (v-mean[c])/std[c]; output NHWC with c_pad channels.  tmp: [batch,3,hin/factor,win] floats */
int e4s_bicubic_down_norm_f32(const float* x, int batch, int hin, int win, int factor, const float* taps,
                              const float* mean, const float* std, float* tmp, float* y, int c_pad, int do_clamp, void* stream);

/* labels u8 [batch,h,w] -> one-hot float [batch,k,h,w] (labelMap2OneHot, utils/torch_utils.py:207-213) */
int e4s_labels_to_onehot_f32(const uint8_t* labels, int batch, int k, int h, int w, float* onehot, void* stream);

/* Style recombination between encoder and generator (swap_comp_style_vector, reference swap_face_fine/swap_face_mask.py:336-367):
 * target/source/out [batch, ncomp, dim] fp32; bit c of comp_mask = component c takes the source's vector.  Components 7 (ears:
 * average), 11 (ear-rings: target), 8 (below-face: average when below_face) and 9 (mouth: target when the source's vector sums to
 * exactly 0, evaluated per sample) follow the reference's fixed rules. */
int e4s_swap_comp_styles_f32(const float* target, const float* source, float* out, int batch, int ncomp, int dim,
                             uint32_t comp_mask, int below_face, void* stream);

/* tensor2im on the device (reference utils/torch_utils.py:64-76), batched: x NCHW [batch,3,h,w] fp32 -> y NHWC [batch,h,w,3] u8:
 * (v+1)/2 when zero_center, clamp [0,1], *255, truncate (the reference's float32 numpy arithmetic: identical bytes) */
int e4s_tensor2im_u8(const float* x, uint8_t* y, int batch, int h, int w, int zero_center, void* stream);

/* TO_TENSOR (+ NORMALIZE) on the device (reference datasets/dataset.py:45 and its callers face_swap_video_pipeline.py:338-346,
 * face_parsing_demo.py:153), batched: x u8 HWC [batch,h,w,3] -> y01 = v/255 and / or ynorm = (v/255 - mean[c]) / std[c], both fp32 NCHW
 * [batch,3,h,w] (either may be NULL).  mean3 / std3 are HOST arrays of 3 floats.  Same fp32 operations as torchvision: identical floats. */
int e4s_im2tensor_f32(const uint8_t* x, float* y01, float* ynorm, int batch, int h, int w, const float* mean3, const float* std3,
                      void* stream);

/* Grey-scale dilation / erosion of NCHW planes (reference utils/morphology.py:23-200): x, out [planes, h, w] fp32; neighborhood
 * [se_h, se_w] = 0 (or the non-flat structuring element) where the kernel is non-zero and -max_val elsewhere; out-of-image taps
 * read border_value (geodesic border: -max_val for dilation, +max_val for erosion).  dilate != 0: max(P + flipped neighborhood),
 * else min(P - neighborhood). */
int e4s_morphology_f32(const float* x, const float* neighborhood, float* out, int64_t planes, int h, int w, int se_h, int se_w,
                       int origin_y, int origin_x, float border_value, int dilate, void* stream);

/* ---- paste-back, NCHW fp32 planes [planes, h, w] ------------------------------------------------------------------
 * e4s_depthwise_conv_f32: out = correlate(x, weight[k][k]) with zero padding k/2 (F.conv2d(x, w, groups=C, padding=r) of SoftErosion,
 *   reference utils/paste_back_tricks.py:32-36); min_with_input != 0: out = min(x, conv(x)) (the iterations - 1 first passes).
 * e4s_soft_erosion_finish_f32: in place, x >= threshold -> 1 (mask byte 1), else x / max(x below threshold) over the WHOLE tensor
 *   (paste_back_tricks.py:38-40); scratch_max: one device float.
 * e4s_pyr_down_f32 / e4s_pyr_up_f32: cv2.pyrDown / cv2.pyrUp arithmetic ([1 4 6 4 1] Gaussian, BORDER_REFLECT_101, output
 *   (h+1)/2 x (w+1)/2 resp. 2h x 2w; multi_band_blending.py:17-19,30-31,47).  round_u8: cv2's uint8 form floor((sum+128)/256) on integer
 *   pixel values.  pyr_up mode 0: out = up(x); 1: out = other - up(x) (Laplacian level); 2: out = up(x) + other (reconstruction).
 * e4s_pyr_blend_f32: out = la*gm + lb*(1-gm) (multi_band_blending.py:40-42); gm [batch, m_channels, h, w], m_channels = channels or 1. */
int e4s_depthwise_conv_f32(const float* x, const float* weight, float* out, int64_t planes, int h, int w, int k, int min_with_input,
                           void* stream);
int e4s_soft_erosion_finish_f32(float* x, uint8_t* mask, int64_t n, float threshold, float* scratch_max, void* stream);
int e4s_pyr_down_f32(const float* x, float* out, int64_t planes, int h, int w, int round_u8, void* stream);
int e4s_pyr_up_f32(const float* x, const float* other, float* out, int64_t planes, int h, int w, int mode, void* stream);
int e4s_pyr_blend_f32(const float* la, const float* lb, const float* gm, float* out, int batch, int channels, int m_channels, int h, int w,
                      void* stream);

/* ---- backward pass of the regional modulated convolution (PTI fine-tuning, reference training/video_swap_ft_coach.py:242-318
 * calls loss.backward() through net.G; the reference differentiates its convolutions through models/stylegan2/op/conv2d_gradfix.py:134-225).
 * Data gradients re-use e4s_conv_tc / e4s_conv_f32 with transposed weights; these entry points cover the rest.  NHWC fp32 throughout.
 *
 * e4s_conv_wgrad_f32: dw[co][ci][ky][kx] (+)= scale * sum_b sum_{y < hl, x < wl} sscale[b*s_stride + ci] *
 *        x[b, y + ky*tx - px, x + kx*tx - px, ci] * g[b, y*sg + ky*tg - pg, x*sg + kx*tg - pg, co]      (out-of-range taps contribute 0)
 *   same-resolution k x k conv (pad k/2): hl x wl = output size, tx = 1, px = k/2, sg = 1, tg = 0, pg = 0
 *   conv_transpose2d(stride 2):          hl x wl = input size,  tx = 0, px = 0,   sg = 2, tg = 1, pg = 0 (g on the (2h+1) x (2w+1) grid)
 *   F.linear:                            kh = kw = 1, hl = rows, wl = 1
 *   sscale may be NULL.  Deterministic (fixed split over pixel chunks, ordered second pass); ws: e4s_conv_wgrad_ws_bytes bytes.
 * e4s_region_scale_f32: out[p, c] = g[p, c] * table[(b*regions + r(p))*c + c] for c < c (table may be NULL), 0 for c <= ch < out_c (channel
 *   padding) and for every pixel whose region differs from select_region (>= 0; -1 keeps all): the reference's `* mask_k` (model.py:395-398).
 * e4s_region_dot_f32: out[b, r, c] = sum over the pixels p of sample b with r(p) == r of a[p,c] * b[p,c] (labels NULL: one region).
 *   fp64 partial sums over fixed 1024-pixel chunks, ordered reduction; ws: e4s_region_dot_ws_bytes bytes.
 * e4s_chan_scale_accum_f32: dx[p, c] (+)= h[p, c] * s[b*s_stride + c]  (s may be NULL: plain copy / accumulate). */
int64_t e4s_conv_wgrad_ws_bytes(int batch, int hl, int wl, int cin, int cout, int kh, int kw);
int e4s_conv_wgrad_f32(const float* x, int64_t x_pitch, int batch, int hx, int wx, int cin, const float* g, int64_t g_pitch, int hg, int wg,
                       int cout, int hl, int wl, int kh, int kw, int tx, int px, int sg, int tg, int pg, const float* sscale,
                       int64_t s_stride, float scale, float* dw, int accumulate, void* ws, void* stream);
int e4s_region_scale_f32(const float* g, int64_t g_pitch, int batch, int h, int w, int c, const float* table, const uint8_t* labels,
                         int regions, int lab_h, int lab_w, int select_region, float* out, int64_t out_pitch, int out_c, void* stream);
int64_t e4s_region_dot_ws_bytes(int batch, int h, int w, int c, int regions);
int e4s_region_dot_f32(const float* a, int64_t a_pitch, const float* b, int64_t b_pitch, int batch, int h, int w, int c,
                       const uint8_t* labels, int regions, int lab_h, int lab_w, float* out, void* ws, void* stream);
int e4s_chan_scale_accum_f32(const float* h, int64_t h_pitch, const float* s, int64_t s_stride, float* dx, int64_t dx_pitch, int batch,
                             int64_t hw, int c, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* E4S_B200_H */

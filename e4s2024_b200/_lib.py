"""ctypes binding of libe4s_b200.so (include/e4s_b200.h) + thin torch-tensor wrappers.

There is NO fallback: if the library cannot be loaded, or a call fails, an exception is raised.
Every wrapper enqueues on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libe4s_b200.so")

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_PRELU, ACT_SIGMOID, ACT_RSQRT_EPS = range(6)
TC_BF16, TC_F16 = 0, 1
CONV_NORMAL, CONV_UP2 = 0, 1

_fp = C.c_void_p
_i32, _i64, _f32 = C.c_int32, C.c_int64, C.c_float


class E4SConv(C.Structure):
    """Mirror of `struct E4SConv` (include/e4s_b200.h); field order is the ABI."""
    _fields_ = [
        ("x", _fp), ("x_pitch", _i64),
        ("batch", _i32), ("hin", _i32), ("win", _i32), ("cin", _i32),
        ("in_shift", _i32), ("in_square", _i32),
        ("w", _fp),
        ("cout", _i32), ("cout_pad", _i32),
        ("kh", _i32), ("kw", _i32), ("stride", _i32), ("pad", _i32), ("mode", _i32),
        ("hout", _i32), ("wout", _i32),
        ("in_mean", _fp), ("in_rstd", _fp),
        ("smod", _fp), ("labels", _fp),
        ("regions", _i32), ("lab_h", _i32), ("lab_w", _i32),
        ("demod", _fp),
        ("pixw", _fp), ("pixw_sb", _i64),
        ("ch_scale", _fp), ("ch_shift", _fp),
        ("noise", _fp), ("noise_w", _fp), ("noise_sb", _i64), ("noise_sc", _i64),
        ("res", _fp), ("res_pitch", _i64), ("res_after_act", _i32),
        ("act", _i32), ("act_slope", _f32), ("act_gain", _f32), ("act_prelu", _fp),
        ("out", _fp), ("out_pitch", _i64), ("accumulate", _i32),
        ("rgb", _fp), ("rgb_w", _fp), ("rgb_smod", _fp), ("rgb_bias", _fp), ("rgb_skip", _fp), ("rgb_fir", _fp),
        ("pred_count", _fp), ("pred_limit", _i32), ("pred_run_if_gt", _i32),
        ("tc_fmt", _i32), ("tc_out_scale", _f32), ("tc_unbias", _f32),
    ]


EXPORTS = [
    "e4s_last_error", "e4s_launch_count", "e4s_device_info", "e4s_sizeof_conv", "e4s_conv_f32", "e4s_conv_f32_batched", "e4s_conv_tc", "e4s_conv_tc_regions", "e4s_conv_tc_splitk_ws_bytes", "e4s_conv_tc_splitk", "e4s_region_tile_jobs", "e4s_upz_build_rows", "e4s_pack_convt_weights_f32", "e4s_conv_tc_upz",
    "e4s_debug_halo_trace", "e4s_debug_halo_flags", "e4s_debug_upz_flags", "e4s_pack_weights_tc_bytes", "e4s_pack_weights_tc", "e4s_pack_weights_tc_fmt", "e4s_pack_conv_weights_f32", "e4s_pack_upconv_weights_f32", "e4s_upfirdn2d_f32", "e4s_bias_act_f32",
    "e4s_bias_act_grad_f32",
    "e4s_noise_bias_act_nhwc_f32", "e4s_nchw_to_nhwc_f32", "e4s_nhwc_to_nchw_f32", "e4s_mask_labels",
    "e4s_torgb_f32", "e4s_chan_stats_ws_bytes", "e4s_chan_stats_f32", "e4s_vec_fc_f32",
    "e4s_residual_combine_f32", "e4s_masked_mean_f32", "e4s_mask_member_bits_u32", "e4s_masked_mean_ws_bytes", "e4s_masked_mean_bits_f32", "e4s_resize_bilinear_nchw_to_nhwc_f32",
    "e4s_resize_bilinear_nhwc_to_nchw_f32", "e4s_maxpool3x3s2_nhwc_f32", "e4s_upsample_argmax_u8",
    "e4s_bicubic_down_norm_f32", "e4s_labels_to_onehot_f32", "e4s_swap_comp_styles_f32", "e4s_tensor2im_u8", "e4s_im2tensor_f32", "e4s_morphology_f32", "e4s_depthwise_conv_f32", "e4s_soft_erosion_finish_f32", "e4s_pyr_down_f32", "e4s_pyr_up_f32", "e4s_pyr_blend_f32",
    "e4s_conv_wgrad_ws_bytes", "e4s_conv_wgrad_f32", "e4s_region_scale_f32", "e4s_region_dot_ws_bytes", "e4s_region_dot_f32", "e4s_chan_scale_accum_f32",
]

_lib = None


class E4SError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise E4SError(f"{LIB_PATH} is missing: run `python -m e4s2024_b200.build` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.e4s_last_error.restype = C.c_char_p
        _lib.e4s_launch_count.restype = C.c_int64
        _lib.e4s_chan_stats_ws_bytes.restype = C.c_int64
        _lib.e4s_pack_weights_tc_bytes.restype = C.c_int64
        _lib.e4s_masked_mean_ws_bytes.restype = C.c_int64
        _lib.e4s_conv_wgrad_ws_bytes.restype = C.c_int64
        _lib.e4s_conv_tc_splitk_ws_bytes.restype = C.c_int64
        _lib.e4s_region_dot_ws_bytes.restype = C.c_int64
        if _lib.e4s_sizeof_conv() != C.sizeof(E4SConv):
            raise E4SError(f"struct E4SConv mismatch: C {_lib.e4s_sizeof_conv()} vs ctypes {C.sizeof(E4SConv)}")
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise E4SError(f"{what} failed ({rc}): {lib().e4s_last_error().decode()}")


def launch_count() -> int:
    return int(lib().e4s_launch_count())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def check_device(t: torch.Tensor):
    """Kernels are enqueued on the CURRENT device's stream: a tensor that lives on another GPU would be dereferenced by the wrong
    device.  One process per GPU with torch.cuda.set_device (bench.py, sharding.py) never trips this."""
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise E4SError(f"tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: wrap the call in "
                       f"`with torch.cuda.device({t.device.index}):` or call torch.cuda.set_device first")


def _req(t: torch.Tensor, dtype=torch.float32, contig: bool = True):
    if not t.is_cuda:
        raise E4SError("e4s2024_b200 kernels need CUDA tensors (no CPU fallback)")
    check_device(t)
    if t.dtype != dtype:
        raise E4SError(f"expected {dtype}, got {t.dtype}")
    if contig and not t.is_contiguous():
        raise E4SError("tensor must be contiguous")
    return t


# ------------------------------------------------------------------------------------------------
# wrappers (one per C entry point)
# ------------------------------------------------------------------------------------------------

def conv(params: E4SConv, tc_weights: Optional[torch.Tensor] = None):
    if tc_weights is not None:
        _check(lib().e4s_conv_tc(C.byref(params), C.c_void_p(tc_weights.data_ptr()), _stream()), "e4s_conv_tc")
    else:
        _check(lib().e4s_conv_f32(C.byref(params), _stream()), "e4s_conv_f32")


def conv_splitk(params: E4SConv, tc_weights: torch.Tensor, ksplit: int):
    """e4s_conv_tc_splitk: the partial-accumulator workspace is allocated here (torch's caching allocator) for the duration of the call."""
    n = int(lib().e4s_conv_tc_splitk_ws_bytes(C.byref(params), int(ksplit)))
    ws = torch.empty(n // 4, dtype=torch.float32, device=tc_weights.device)
    _check(lib().e4s_conv_tc_splitk(C.byref(params), C.c_void_p(tc_weights.data_ptr()), int(ksplit), _fp(ws.data_ptr()), _stream()), "e4s_conv_tc_splitk")


def conv_regions(params: E4SConv, tc_weights: torch.Tensor, jobs: torch.Tensor, count: torch.Tensor, count_host: int):
    _check(lib().e4s_conv_tc_regions(C.byref(params), C.c_void_p(tc_weights.data_ptr()), C.c_void_p(jobs.data_ptr()),
                                     C.c_void_p(count.data_ptr()), int(count_host), _stream()), "e4s_conv_tc_regions")


def region_tile_jobs(labels: torch.Tensor, hout: int, wout: int, up2: bool, regions: int, count_slot: torch.Tensor):
    """-> jobs int32 [max_jobs, 4]; count_slot: zeroed int32[1] view that receives the job count."""
    b, lh, lw = labels.shape
    gh, gw = (hout // 2, wout // 2) if up2 else (hout, wout)
    max_jobs = b * (gh // 16) * (gw // 8) * min(regions, 32)
    jobs = torch.empty(max_jobs, 4, dtype=torch.int32, device=labels.device)
    _check(lib().e4s_region_tile_jobs(_fp(labels.data_ptr()), b, lh, lw, hout, wout, int(up2), _fp(jobs.data_ptr()),
                                      _fp(count_slot.data_ptr()), max_jobs, _stream()), "e4s_region_tile_jobs")
    return jobs


def upz_build_rows(labels: torch.Tensor, hin: int, win: int, max_rows: int, count_slot: torch.Tensor):
    """-> (cells int32 [B, hin+1, win+1, 2], rows int32 [max_rows, 2]); count_slot: zeroed int32[1] view that receives the row count."""
    b, lh, lw = labels.shape
    cells = torch.empty(b, hin + 1, win + 1, 2, dtype=torch.int32, device=labels.device)
    rows = torch.empty(max_rows, 2, dtype=torch.int32, device=labels.device)
    _check(lib().e4s_upz_build_rows(_fp(labels.data_ptr()), b, lh, lw, hin, win, _fp(cells.data_ptr()), _fp(rows.data_ptr()),
                                    _fp(count_slot.data_ptr()), max_rows, _stream()), "e4s_upz_build_rows")
    return cells, rows


def pack_convt_weights(w: torch.Tensor, cout_pad: int, scale: float = 1.0) -> torch.Tensor:
    """w [Co,Ci,3,3] -> [9, Ci, cout_pad] fp32: the taps of conv_transpose2d as [Ci x Co] matrices in conv_tc_upz's consumption order."""
    _req(w)
    co, ci = w.shape[:2]
    out = torch.empty(9, ci, cout_pad, device=w.device, dtype=torch.float32)
    _check(lib().e4s_pack_convt_weights_f32(_fp(w.data_ptr()), _fp(out.data_ptr()), co, ci, cout_pad, _f32(scale), _stream()),
           "e4s_pack_convt_weights_f32")
    return out


def conv_upz(params: E4SConv, tc9: torch.Tensor, fir: torch.Tensor, cells: Optional[torch.Tensor], rows: Optional[torch.Tensor],
             count: Optional[torch.Tensor], max_rows: int, z: torch.Tensor):
    """cells / rows / count None: the direct (un-masked) form, max_rows = batch * (hin+1) * (win+1)."""
    _check(lib().e4s_conv_tc_upz(C.byref(params), _fp(tc9.data_ptr()), _fp(fir.data_ptr()), _fp(_p(cells)), _fp(_p(rows)),
                                 _fp(_p(count)), int(max_rows), _fp(z.data_ptr()), _stream()), "e4s_conv_tc_upz")


def conv_batched(params_list):
    """Launch a list of E4SConv problems (mode NORMAL, fp32 engine) as one batched kernel."""
    n = len(params_list)
    arr = (E4SConv * n)(*params_list)
    _check(lib().e4s_conv_f32_batched(arr, n, _stream()), "e4s_conv_f32_batched")


def pack_weights_tc(w_f32: torch.Tensor, phases: int, k: int, cin: int, cout: int, cout_pad: int, fmt: int = TC_BF16,
                    scale: float = 1.0) -> torch.Tensor:
    nbytes = int(lib().e4s_pack_weights_tc_bytes(phases, k, cout))
    if nbytes <= 0:
        raise E4SError("e4s_pack_weights_tc_bytes: unsupported shape")
    out = torch.empty(nbytes, dtype=torch.uint8, device=w_f32.device)
    _check(lib().e4s_pack_weights_tc_fmt(C.c_void_p(w_f32.data_ptr()), phases, k, cin, cout, cout_pad, int(fmt), _f32(scale),
                                         C.c_void_p(out.data_ptr()), _stream()), "e4s_pack_weights_tc_fmt")
    return out


def pack_conv_weights(w: torch.Tensor, cin_pad: int, cout_pad: int, scale: float = 1.0, sumsq: bool = False) -> torch.Tensor:
    """w [Co,Ci,kh,kw] -> [1, kh*kw*cin_pad, cout_pad] fp32 (sumsq: [1, cin_pad, cout_pad] = sum_taps (scale*w)^2)."""
    _req(w)
    co, ci, kh, kw = w.shape
    out = torch.empty(1, (1 if sumsq else kh * kw) * cin_pad, cout_pad, device=w.device, dtype=torch.float32)
    _check(lib().e4s_pack_conv_weights_f32(_fp(w.data_ptr()), _fp(out.data_ptr()), co, ci, kh, kw, cin_pad, cout_pad, _f32(scale), int(sumsq),
                                           _stream()), "e4s_pack_conv_weights_f32")
    return out


def pack_upconv_weights(w: torch.Tensor, fir: torch.Tensor, cout_pad: int, scale: float = 1.0) -> torch.Tensor:
    """w [Co,Ci,3,3], fir [4,4] -> [4, 9*Ci, cout_pad] fp32 phase filters."""
    _req(w), _req(fir)
    co, ci = w.shape[:2]
    out = torch.empty(4, 9 * ci, cout_pad, device=w.device, dtype=torch.float32)
    _check(lib().e4s_pack_upconv_weights_f32(_fp(w.data_ptr()), _fp(fir.data_ptr()), _fp(out.data_ptr()), co, ci, cout_pad, _f32(scale),
                                             _stream()), "e4s_pack_upconv_weights_f32")
    return out


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int, down: int, pad0: int, pad1: int) -> torch.Tensor:
    _req(x), _req(kernel)
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    oh = (h * up + pad0 + pad1 - kh + down) // down
    ow = (w * up + pad0 + pad1 - kw + down) // down
    out = torch.empty(b, c, max(oh, 0), max(ow, 0), device=x.device, dtype=x.dtype)
    _check(lib().e4s_upfirdn2d_f32(_fp(x.data_ptr()), _fp(kernel.data_ptr()), _fp(out.data_ptr()), C.c_int64(b * c), h, w,
                                   kh, kw, up, up, down, down, pad0, pad1, pad0, pad1, _stream()), "e4s_upfirdn2d_f32")
    return out


def upfirdn2d_general(x: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                      pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """The full parameter set of the reference op (upfirdn2d.cpp:12-23): separate x/y factors and the four pads."""
    _req(x), _req(kernel)
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    oh = (h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y
    ow = (w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    out = torch.empty(b, c, max(oh, 0), max(ow, 0), device=x.device, dtype=x.dtype)
    _check(lib().e4s_upfirdn2d_f32(_fp(x.data_ptr()), _fp(kernel.data_ptr()), _fp(out.data_ptr()), C.c_int64(b * c), h, w,
                                   kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, _stream()), "e4s_upfirdn2d_f32")
    return out


def bias_act_grad(g: torch.Tensor, bias: Optional[torch.Tensor], ref: torch.Tensor, slope: float, scale: float) -> torch.Tensor:
    _req(g), _req(ref)
    y = torch.empty_like(g)
    channels = g.shape[1] if g.ndim > 1 else 1
    inner = 1
    for d in g.shape[2:]:
        inner *= d
    _check(lib().e4s_bias_act_grad_f32(_fp(g.data_ptr()), _fp(_p(bias)), _fp(ref.data_ptr()), _fp(y.data_ptr()), C.c_int64(g.numel()),
                                       C.c_int64(inner), channels, _f32(slope), _f32(scale), _stream()), "e4s_bias_act_grad_f32")
    return y


def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], slope: float, scale: float) -> torch.Tensor:
    _req(x)
    y = torch.empty_like(x)
    channels = x.shape[1] if x.ndim > 1 else 1
    inner = 1
    for d in x.shape[2:]:
        inner *= d
    _check(lib().e4s_bias_act_f32(_fp(x.data_ptr()), _fp(_p(bias)), _fp(y.data_ptr()), C.c_int64(x.numel()), C.c_int64(inner),
                                  channels, _f32(slope), _f32(scale), _stream()), "e4s_bias_act_f32")
    return y


def noise_bias_act_nhwc(x, noise, noise_w, noise_sb, noise_sc, bias, slope, scale):
    b, h, w, c = x.shape
    _check(lib().e4s_noise_bias_act_nhwc_f32(_fp(x.data_ptr()), b, h, w, c, _fp(_p(noise)), _fp(_p(noise_w)), C.c_int64(noise_sb),
                                             C.c_int64(noise_sc), _fp(_p(bias)), _f32(slope), _f32(scale), _stream()),
           "e4s_noise_bias_act_nhwc_f32")


def nchw_to_nhwc(x: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    _req(x)
    b, c, h, w = x.shape
    c_pad = c if c_pad is None else c_pad
    y = torch.empty(b, h, w, c_pad, device=x.device, dtype=x.dtype)
    _check(lib().e4s_nchw_to_nhwc_f32(_fp(x.data_ptr()), _fp(y.data_ptr()), b, c, h, w, c_pad, _stream()), "e4s_nchw_to_nhwc_f32")
    return y


def nhwc_to_nchw(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    """x: [B,H,W,P] contiguous; the first c channels are converted."""
    _req(x)
    b, h, w, pitch = x.shape
    c = pitch if c is None else c
    y = torch.empty(b, c, h, w, device=x.device, dtype=x.dtype)
    _check(lib().e4s_nhwc_to_nchw_f32(_fp(x.data_ptr()), C.c_int64(pitch), _fp(y.data_ptr()), b, c, h, w, _stream()),
           "e4s_nhwc_to_nchw_f32")
    return y


def mask_labels(mask: torch.Tensor, flags: Optional[torch.Tensor] = None):
    """-> (labels u8 [B,H,W], flags int32[1]); flags[0] > 0 means the mask is not one-hot/empty per pixel."""
    _req(mask)
    b, k, h, w = mask.shape
    labels = torch.empty(b, h, w, device=mask.device, dtype=torch.uint8)
    if flags is None:
        flags = torch.zeros(1, device=mask.device, dtype=torch.int32)
    _check(lib().e4s_mask_labels(_fp(mask.data_ptr()), b, k, h, w, _fp(labels.data_ptr()), _fp(flags.data_ptr()), _stream()),
           "e4s_mask_labels")
    return labels, flags


def labels_to_onehot(labels: torch.Tensor, k: int) -> torch.Tensor:
    _req(labels, torch.uint8)
    b, h, w = labels.shape
    out = torch.empty(b, k, h, w, device=labels.device, dtype=torch.float32)
    _check(lib().e4s_labels_to_onehot_f32(_fp(labels.data_ptr()), b, k, h, w, _fp(out.data_ptr()), _stream()), "e4s_labels_to_onehot_f32")
    return out


def tensor2im_u8(x: torch.Tensor, zero_center: bool = True) -> torch.Tensor:
    """[B,3,H,W] fp32 -> [B,H,W,3] uint8 on the device."""
    _req(x)
    b, c, h, w = x.shape
    if c != 3:
        raise E4SError("tensor2im_u8 expects 3 channels")
    y = torch.empty(b, h, w, 3, device=x.device, dtype=torch.uint8)
    _check(lib().e4s_tensor2im_u8(_fp(x.data_ptr()), _fp(y.data_ptr()), b, h, w, int(zero_center), _stream()), "e4s_tensor2im_u8")
    return y


def im2tensor(x: torch.Tensor, want01: bool = True, want_norm: bool = True, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """u8 HWC [B,H,W,3] on the device -> (x01 [B,3,H,W] or None, xnorm [B,3,H,W] or None), fp32 (TO_TENSOR / NORMALIZE)."""
    _req(x, torch.uint8)
    b, h, w, c = x.shape
    if c != 3:
        raise E4SError("im2tensor expects [B,H,W,3]")
    y01 = torch.empty(b, 3, h, w, device=x.device, dtype=torch.float32) if want01 else None
    yn = torch.empty(b, 3, h, w, device=x.device, dtype=torch.float32) if want_norm else None
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _check(lib().e4s_im2tensor_f32(_fp(x.data_ptr()), _fp(_p(y01)), _fp(_p(yn)), b, h, w, m3, s3, _stream()), "e4s_im2tensor_f32")
    return y01, yn


def morphology(x: torch.Tensor, neighborhood: torch.Tensor, origin, border_value: float, dilate: bool) -> torch.Tensor:
    """x [B,C,H,W] fp32, neighborhood [se_h,se_w] fp32 (see e4s_morphology_f32) -> [B,C,H,W]."""
    _req(x), _req(neighborhood)
    b, c, h, w = x.shape
    se_h, se_w = neighborhood.shape
    out = torch.empty_like(x)
    _check(lib().e4s_morphology_f32(_fp(x.data_ptr()), _fp(neighborhood.data_ptr()), _fp(out.data_ptr()), C.c_int64(b * c), h, w, se_h, se_w,
                                    int(origin[0]), int(origin[1]), C.c_float(border_value), int(dilate), _stream()), "e4s_morphology_f32")
    return out


def depthwise_conv(x: torch.Tensor, weight: torch.Tensor, min_with_input: bool = False) -> torch.Tensor:
    """x [B,C,H,W], weight [k,k] (odd k): depthwise correlation with zero padding k//2 (optionally min(x, conv(x)))."""
    _req(x), _req(weight)
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    _check(lib().e4s_depthwise_conv_f32(_fp(x.data_ptr()), _fp(weight.data_ptr()), _fp(out.data_ptr()), C.c_int64(b * c), h, w, weight.shape[0],
                                        int(min_with_input), _stream()), "e4s_depthwise_conv_f32")
    return out


def soft_erosion_finish(x: torch.Tensor, threshold: float):
    """in place on x [..] fp32 -> (x, mask bool): x >= thr -> 1, else x / max(x below thr)."""
    _req(x)
    mask = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    scratch = torch.empty(1, device=x.device, dtype=torch.float32)
    _check(lib().e4s_soft_erosion_finish_f32(_fp(x.data_ptr()), _fp(mask.data_ptr()), C.c_int64(x.numel()), _f32(threshold), _fp(scratch.data_ptr()),
                                             _stream()), "e4s_soft_erosion_finish_f32")
    return x, mask.bool()


def pyr_down(x: torch.Tensor, round_u8: bool = False) -> torch.Tensor:
    _req(x)
    b, c, h, w = x.shape
    out = torch.empty(b, c, (h + 1) // 2, (w + 1) // 2, device=x.device, dtype=torch.float32)
    _check(lib().e4s_pyr_down_f32(_fp(x.data_ptr()), _fp(out.data_ptr()), C.c_int64(b * c), h, w, int(round_u8), _stream()), "e4s_pyr_down_f32")
    return out


def pyr_up(x: torch.Tensor, other: Optional[torch.Tensor] = None, mode: int = 0) -> torch.Tensor:
    """mode 0: up(x); 1: other - up(x); 2: up(x) + other.  other [B,C,2H,2W]."""
    _req(x)
    b, c, h, w = x.shape
    if other is not None:
        _req(other)
        if tuple(other.shape) != (b, c, 2 * h, 2 * w):
            raise E4SError(f"pyr_up: other must be {(b, c, 2 * h, 2 * w)}, got {tuple(other.shape)}")
    out = torch.empty(b, c, 2 * h, 2 * w, device=x.device, dtype=torch.float32)
    _check(lib().e4s_pyr_up_f32(_fp(x.data_ptr()), _fp(_p(other)), _fp(out.data_ptr()), C.c_int64(b * c), h, w, mode, _stream()), "e4s_pyr_up_f32")
    return out


def pyr_blend(la: torch.Tensor, lb: torch.Tensor, gm: torch.Tensor) -> torch.Tensor:
    _req(la), _req(lb), _req(gm)
    b, c, h, w = la.shape
    out = torch.empty_like(la)
    _check(lib().e4s_pyr_blend_f32(_fp(la.data_ptr()), _fp(lb.data_ptr()), _fp(gm.data_ptr()), _fp(out.data_ptr()), b, c, gm.shape[1], h, w, _stream()),
           "e4s_pyr_blend_f32")
    return out


def swap_comp_styles(target: torch.Tensor, source: torch.Tensor, comp_mask: int, below_face: bool) -> torch.Tensor:
    _req(target), _req(source)
    b, k, d = target.shape
    out = torch.empty_like(target)
    _check(lib().e4s_swap_comp_styles_f32(_fp(target.data_ptr()), _fp(source.data_ptr()), _fp(out.data_ptr()), b, k, d,
                                          C.c_uint32(comp_mask), int(below_face), _stream()), "e4s_swap_comp_styles_f32")
    return out


def torgb(x_nhwc, cin, smod, wrgb, labels, regions, lab_hw, pixw, pixw_sb, bias, skip, fir, rgb, accumulate):
    b, h, w, pitch = x_nhwc.shape
    lh, lw = lab_hw
    _check(lib().e4s_torgb_f32(_fp(x_nhwc.data_ptr()), C.c_int64(pitch), b, h, w, cin, _fp(smod.data_ptr()), _fp(wrgb.data_ptr()),
                               _fp(_p(labels)), regions, lh, lw, _fp(pixw), C.c_int64(pixw_sb), _fp(_p(bias)), _fp(_p(skip)),
                               _fp(_p(fir)), _fp(rgb.data_ptr()), int(accumulate), _stream()), "e4s_torgb_f32")


_ws_cache = {}


def _stats_ws(device, batch, c) -> torch.Tensor:
    n = int(lib().e4s_chan_stats_ws_bytes(batch, c))
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.zeros(max(n, 1 << 20), dtype=torch.uint8, device=device)      # zero-filled: the kernel's arrival counters (e4s_b200.h)
        _ws_cache[key] = ws
    return ws


def chan_stats(x_nhwc: torch.Tensor, c: int, eps: float = 1e-5, want_rstd: bool = True):
    """x [B,H,W,P]: (mean [B,c], rstd [B,c] or None) over H*W of the first c channels."""
    b, h, w, pitch = x_nhwc.shape
    mean = torch.empty(b, c, device=x_nhwc.device, dtype=torch.float32)
    rstd = torch.empty(b, c, device=x_nhwc.device, dtype=torch.float32) if want_rstd else None
    ws = _stats_ws(x_nhwc.device, b, c)
    _check(lib().e4s_chan_stats_f32(_fp(x_nhwc.data_ptr()), C.c_int64(pitch), b, h * w, c, _f32(eps), _fp(mean.data_ptr()),
                                    _fp(_p(rstd)), _fp(ws.data_ptr()), _stream()), "e4s_chan_stats_f32")
    return mean, rstd


def vec_fc(x: torch.Tensor, w: torch.Tensor, scale, shift, act: int) -> torch.Tensor:
    """x [B,cin] (row stride x.stride(0)), w [cout,cin] -> [B,cout]."""
    b, cin = x.shape
    cout = w.shape[0]
    y = torch.empty(b, cout, device=x.device, dtype=torch.float32)
    _check(lib().e4s_vec_fc_f32(_fp(x.data_ptr()), C.c_int64(x.stride(0)), _fp(w.data_ptr()), _fp(_p(scale)), _fp(_p(shift)),
                                _fp(y.data_ptr()), b, cin, cout, act, _stream()), "e4s_vec_fc_f32")
    return y


def residual_combine(a, c, a_stats=None, gate=None, gate_plus_one=False, r=None, r_sub=1, r_stats=None, relu=False,
                     prelu=None, out=None):
    """a [B,H,W,Pa]; r [B,Hr,Wr,Pr] or None -> out [B,H,W,c] (see e4s_residual_combine_f32)."""
    b, h, w, pa = a.shape
    if out is None:
        out = torch.empty(b, h, w, c, device=a.device, dtype=torch.float32)
    am, ar = a_stats if a_stats is not None else (None, None)
    rm, rr = r_stats if r_stats is not None else (None, None)
    rh, rw, pr = (r.shape[1], r.shape[2], r.shape[3]) if r is not None else (0, 0, 0)
    _check(lib().e4s_residual_combine_f32(_fp(a.data_ptr()), C.c_int64(pa), _fp(_p(am)), _fp(_p(ar)), _fp(_p(gate)),
                                          int(gate_plus_one), _fp(_p(r)), C.c_int64(pr), rh, rw, r_sub, _fp(_p(rm)), _fp(_p(rr)),
                                          int(relu), _fp(_p(prelu)), _fp(out.data_ptr()), C.c_int64(out.shape[3]), b, h, w, c, _stream()),
           "e4s_residual_combine_f32")
    return out


def masked_mean(feat_nhwc, c, mask, codes, c_off):
    """codes [B,K,D] gets codes[:, :, c_off:c_off+c] = per-region mean of feat."""
    b, h, w, pitch = feat_nhwc.shape
    _, k, mh, mw = mask.shape
    _check(lib().e4s_masked_mean_f32(_fp(feat_nhwc.data_ptr()), C.c_int64(pitch), b, h, w, c, _fp(mask.data_ptr()), k, mh, mw,
                                     _fp(codes.data_ptr()), C.c_int64(codes.stride(0)), C.c_int64(codes.stride(1)), c_off,
                                     _stream()), "e4s_masked_mean_f32")


def mask_member_bits(mask: torch.Tensor) -> torch.Tensor:
    """mask [B,K,H,W] float (K <= 32) -> int32 [B,H,W]: bit j set where mask[b,j] != 0 (region membership)."""
    _req(mask)
    b, k, h, w = mask.shape
    bits = torch.empty(b, h, w, device=mask.device, dtype=torch.int32)
    _check(lib().e4s_mask_member_bits_u32(_fp(mask.data_ptr()), b, k, h, w, _fp(bits.data_ptr()), _stream()), "e4s_mask_member_bits_u32")
    return bits


def masked_mean_bits(feat_nhwc, c, bits, k, codes, c_off):
    """codes[:, :, c_off:c_off+c] = per-region mean of feat over the pixels whose membership word has the region's bit set."""
    b, h, w, pitch = feat_nhwc.shape
    _, mh, mw = bits.shape
    n = int(lib().e4s_masked_mean_ws_bytes(b, c, k))
    key = ("mm", feat_nhwc.device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 20), dtype=torch.uint8, device=feat_nhwc.device)
        _ws_cache[key] = ws
    _check(lib().e4s_masked_mean_bits_f32(_fp(feat_nhwc.data_ptr()), C.c_int64(pitch), b, h, w, c, _fp(bits.data_ptr()), k, mh, mw,
                                          _fp(codes.data_ptr()), C.c_int64(codes.stride(0)), C.c_int64(codes.stride(1)), c_off,
                                          _fp(ws.data_ptr()), _stream()), "e4s_masked_mean_bits_f32")


def resize_bilinear_nchw_to_nhwc(x, hout, wout, c_pad, align_corners=False):
    _req(x)
    b, c, h, w = x.shape
    y = torch.empty(b, hout, wout, c_pad, device=x.device, dtype=torch.float32)
    _check(lib().e4s_resize_bilinear_nchw_to_nhwc_f32(_fp(x.data_ptr()), b, c, h, w, _fp(y.data_ptr()), hout, wout, c_pad,
                                                      int(align_corners), _stream()), "e4s_resize_bilinear_nchw_to_nhwc_f32")
    return y


def resize_bilinear_nhwc_to_nchw(x_nhwc, c, hout, wout, align_corners=True):
    b, h, w, pitch = x_nhwc.shape
    y = torch.empty(b, c, hout, wout, device=x_nhwc.device, dtype=torch.float32)
    _check(lib().e4s_resize_bilinear_nhwc_to_nchw_f32(_fp(x_nhwc.data_ptr()), C.c_int64(pitch), b, c, h, w, _fp(y.data_ptr()), hout,
                                                      wout, int(align_corners), _stream()), "e4s_resize_bilinear_nhwc_to_nchw_f32")
    return y


def maxpool3x3s2(x_nhwc):
    b, h, w, c = x_nhwc.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.empty(b, ho, wo, c, device=x_nhwc.device, dtype=torch.float32)
    _check(lib().e4s_maxpool3x3s2_nhwc_f32(_fp(x_nhwc.data_ptr()), b, h, w, c, _fp(y.data_ptr()), _stream()), "e4s_maxpool3x3s2_nhwc_f32")
    return y


def upsample_argmax(logits_nhwc, c, hout, wout, lut=None):
    b, h, w, pitch = logits_nhwc.shape
    labels = torch.empty(b, hout, wout, device=logits_nhwc.device, dtype=torch.uint8)
    _check(lib().e4s_upsample_argmax_u8(_fp(logits_nhwc.data_ptr()), C.c_int64(pitch), b, c, h, w, hout, wout, _fp(_p(lut)),
                                        _fp(labels.data_ptr()), _stream()), "e4s_upsample_argmax_u8")
    return labels


def bicubic_down_norm(x, factor, taps, mean, std, c_pad, clamp=True):
    _req(x)
    b, c, h, w = x.shape
    if c != 3:
        raise E4SError("bicubic_down_norm expects 3 channels")
    tmp = torch.empty(b, 3, h // factor, w, device=x.device, dtype=torch.float32)
    y = torch.empty(b, h // factor, w // factor, c_pad, device=x.device, dtype=torch.float32)
    _check(lib().e4s_bicubic_down_norm_f32(_fp(x.data_ptr()), b, h, w, factor, _fp(taps.data_ptr()), _fp(mean.data_ptr()),
                                           _fp(std.data_ptr()), _fp(tmp.data_ptr()), _fp(y.data_ptr()), c_pad, int(clamp),
                                           _stream()),
           "e4s_bicubic_down_norm_f32")
    return y


# ---- backward pass ---------------------------------------------------------------

def _ws(tag: str, device, nbytes: int) -> torch.Tensor:
    key = (tag, device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def conv_wgrad(x: torch.Tensor, cin: int, g: torch.Tensor, cout: int, loop_hw, kh: int, kw: int, geom, sscale: Optional[torch.Tensor],
               scale: float, dw: torch.Tensor, accumulate: bool):
    """x [B,Hx,Wx,Px], g [B,Hg,Wg,Pg] NHWC (first cin / cout channels); geom = (tx, px, sg, tg, pg); sscale [B, >=cin] (row stride
    sscale.stride(0)) or None; dw [cout, cin, kh, kw] contiguous (see e4s_conv_wgrad_f32)."""
    b, hx, wx, pxp = x.shape
    _, hg, wg, pgp = g.shape
    hl, wl = loop_hw
    tx, px, sg, tg, pg = geom
    assert dw.is_contiguous() and tuple(dw.shape) == (cout, cin, kh, kw), (dw.shape, (cout, cin, kh, kw))
    ws = _ws("wgrad", x.device, int(lib().e4s_conv_wgrad_ws_bytes(b, hl, wl, cin, cout, kh, kw)))
    _check(lib().e4s_conv_wgrad_f32(_fp(x.data_ptr()), C.c_int64(pxp), b, hx, wx, cin, _fp(g.data_ptr()), C.c_int64(pgp), hg, wg, cout, hl, wl,
                                    kh, kw, tx, px, sg, tg, pg, _fp(_p(sscale)), C.c_int64(0 if sscale is None else sscale.stride(0)), _f32(scale),
                                    _fp(dw.data_ptr()), int(accumulate), _fp(ws.data_ptr()), _stream()), "e4s_conv_wgrad_f32")


def region_scale(g: torch.Tensor, c: int, table: Optional[torch.Tensor], labels: Optional[torch.Tensor], regions: int, select: int = -1,
                 out_c: Optional[int] = None) -> torch.Tensor:
    """g [B,H,W,P] -> [B,H,W,out_c]: g * table[b, r(p)] on the pixels of region `select` (-1: all), zero elsewhere / in the channel padding."""
    b, h, w, pitch = g.shape
    out_c = (c + 3) // 4 * 4 if out_c is None else out_c
    out = torch.empty(b, h, w, out_c, device=g.device, dtype=torch.float32)
    lh, lw = (labels.shape[1], labels.shape[2]) if labels is not None else (0, 0)
    _check(lib().e4s_region_scale_f32(_fp(g.data_ptr()), C.c_int64(pitch), b, h, w, c, _fp(_p(table)), _fp(_p(labels)), regions, lh, lw, int(select),
                                      _fp(out.data_ptr()), C.c_int64(out_c), out_c, _stream()), "e4s_region_scale_f32")
    return out


def region_dot(a: torch.Tensor, b_: torch.Tensor, c: int, labels: Optional[torch.Tensor], regions: int) -> torch.Tensor:
    """a, b_ [B,H,W,P*] -> [B, regions, c]: per-region sums of a*b (labels None: regions must be 1)."""
    b, h, w, pa = a.shape
    out = torch.empty(b, regions, c, device=a.device, dtype=torch.float32)
    lh, lw = (labels.shape[1], labels.shape[2]) if labels is not None else (0, 0)
    ws = _ws("rdot", a.device, int(lib().e4s_region_dot_ws_bytes(b, h, w, c, regions)))
    _check(lib().e4s_region_dot_f32(_fp(a.data_ptr()), C.c_int64(pa), _fp(b_.data_ptr()), C.c_int64(b_.shape[3]), b, h, w, c, _fp(_p(labels)), regions,
                                    lh, lw, _fp(out.data_ptr()), _fp(ws.data_ptr()), _stream()), "e4s_region_dot_f32")
    return out


def chan_scale_accum(h: torch.Tensor, c: int, s: Optional[torch.Tensor], dx: torch.Tensor, accumulate: bool):
    """dx[b,y,x,:c] (+)= h[b,y,x,:c] * s[b,:c]; s [B, >=c] with row stride s.stride(0), or None."""
    b, hh, ww, ph = h.shape
    _check(lib().e4s_chan_scale_accum_f32(_fp(h.data_ptr()), C.c_int64(ph), _fp(_p(s)), C.c_int64(0 if s is None else s.stride(0)), _fp(dx.data_ptr()),
                                          C.c_int64(dx.shape[3]), b, C.c_int64(hh * ww), c, int(accumulate), _stream()), "e4s_chan_scale_accum_f32")

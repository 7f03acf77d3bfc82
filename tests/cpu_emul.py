"""TEST INFRASTRUCTURE: a slow torch/CPU emulation of the C-ABI entry points (include/e4s_b200.h).

It lets the `-m "not gpu"` suite execute the HOST logic of the drop-in modules (weight packing,
struct E4SConv filling, pitches / offsets / channel windows, region tables, layer ordering) against
the oracle without a GPU.  It is an executable restatement of the header's semantics, reads the same
raw pointers the CUDA kernels would, and is never importable from the product package.
"""
from __future__ import annotations

import ctypes as C
import math
from contextlib import contextmanager

import numpy as np
import torch
import torch.nn.functional as F

from e4s2024_b200 import _lib as L


def _f32(ptr: int, n: int) -> torch.Tensor:
    """float32 view of `n` floats at raw address `ptr` (memory owned by a live torch tensor)."""
    arr = np.ctypeslib.as_array((C.c_float * n).from_address(ptr))
    return torch.from_numpy(arr)


def _u8(ptr: int, n: int) -> torch.Tensor:
    return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)))


def _nearest(dst: torch.Tensor, in_size: int, out_size: int) -> torch.Tensor:
    if in_size == out_size:
        return dst
    scale = np.float32(in_size) / np.float32(out_size)
    s = torch.floor(dst.to(torch.float32) * float(scale)).long()
    return s.clamp(max=in_size - 1)


def _act(v, act, slope, gain, prelu):
    if act == L.ACT_LRELU:
        return torch.where(v < 0, v * slope, v) * gain
    if act == L.ACT_RELU:
        return v.clamp(min=0)
    if act == L.ACT_PRELU:
        return torch.where(v < 0, v * prelu, v)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(v)
    if act == L.ACT_RSQRT_EPS:
        return torch.rsqrt(v + slope)
    return v


def conv(p: L.E4SConv, tc_weights=None):
    B, Hin, Win, Cin = p.batch, p.hin, p.win, p.cin
    Ho, Wo, Co, Cp = p.hout, p.wout, p.cout, p.cout_pad
    assert Cin % 8 == 0 and Cp % 4 == 0 and p.x_pitch % 4 == 0
    taps = p.kh * p.kw
    K = taps * Cin
    up = p.mode == L.CONV_UP2
    phases = 4 if up else 1
    x = _f32(p.x, (B * Hin * Win - 1) * p.x_pitch + Cin)
    w = _f32(p.w, phases * K * Cp).reshape(phases, K, Cp)
    out = _f32(p.out, (B * Ho * Wo - 1) * p.out_pitch + Co)
    labels = _u8(p.labels, B * p.lab_h * p.lab_w).reshape(B, p.lab_h, p.lab_w).long() if p.labels else None
    for ph in range(phases):
        py, px = ph >> 1, ph & 1
        if up:
            a, bb = torch.meshgrid(torch.arange(Hin), torch.arange(Win), indexing="ij")
            oy, ox = 2 * a + py, 2 * bb + px
        else:
            oy, ox = torch.meshgrid(torch.arange(Ho), torch.arange(Wo), indexing="ij")
        oy, ox = oy.reshape(-1), ox.reshape(-1)
        npx = oy.numel()
        bidx = torch.arange(B).repeat_interleave(npx)
        oyb, oxb = oy.repeat(B), ox.repeat(B)
        sy, sx = _nearest(oyb, p.lab_h, Ho), _nearest(oxb, p.lab_w, Wo)
        reg = labels[bidx, sy, sx] if labels is not None else torch.zeros_like(bidx)
        A = torch.zeros(B * npx, taps, Cin)
        hv, wv = Hin << p.in_shift, Win << p.in_shift
        for t in range(taps):
            if up:
                iy = (oyb - py) // 2 - 1 + t // 3
                ix = (oxb - px) // 2 - 1 + t % 3
            else:
                iy = oyb * p.stride - p.pad + t // p.kw
                ix = oxb * p.stride - p.pad + t % p.kw
            ok = (iy >= 0) & (iy < hv) & (ix >= 0) & (ix < wv)
            iy2, ix2 = (iy.clamp(0, hv - 1) >> p.in_shift), (ix.clamp(0, wv - 1) >> p.in_shift)
            base = ((bidx * Hin + iy2) * Win + ix2) * p.x_pitch
            v = x[base[:, None] + torch.arange(Cin)[None, :]]
            if p.in_mean:
                mean = _f32(p.in_mean, B * Cin).reshape(B, Cin)
                rstd = _f32(p.in_rstd, B * Cin).reshape(B, Cin)
                v = (v - mean[bidx]) * rstd[bidx]
            if p.in_square:
                v = v * v
            if p.smod:
                sm = _f32(p.smod, ((B - 1) * p.regions + int(reg.max()) + 1) * Cin).reshape(-1, Cin)
                v = v * sm[bidx * p.regions + reg]
            A[:, t] = torch.where(ok[:, None], v, torch.zeros_like(v))
        acc = A.reshape(B * npx, K) @ w[ph][:, :Co]
        if p.demod:
            d = _f32(p.demod, ((B - 1) * p.regions + int(reg.max()) + 1) * Co).reshape(-1, Co)
            acc = acc * d[bidx * p.regions + reg]
        if p.pixw:
            span = (B - 1) * p.pixw_sb + p.lab_h * p.lab_w
            acc = acc * _f32(p.pixw, span)[bidx * p.pixw_sb + sy * p.lab_w + sx][:, None]
        if p.ch_scale:
            acc = acc * _f32(p.ch_scale, Co)[None]
        if p.noise:
            nz = _f32(p.noise, (B - 1) * p.noise_sb + (Co - 1) * p.noise_sc + Ho * Wo)
            idx = bidx[:, None] * p.noise_sb + torch.arange(Co)[None] * p.noise_sc + (oyb * Wo + oxb)[:, None]
            acc = acc + _f32(p.noise_w, 1)[0] * nz[idx]
        if p.ch_shift:
            acc = acc + _f32(p.ch_shift, Co)[None]
        pix = (bidx * Ho + oyb) * Wo + oxb
        res = None
        if p.res:
            rr = _f32(p.res, (B * Ho * Wo - 1) * p.res_pitch + Co)
            res = rr[pix[:, None] * p.res_pitch + torch.arange(Co)[None]]
            if not p.res_after_act:
                acc = acc + res
        prelu = _f32(p.act_prelu, Co)[None] if p.act_prelu else None
        acc = _act(acc, p.act, p.act_slope, p.act_gain, prelu)
        if res is not None and p.res_after_act:
            acc = acc + res
        oidx = pix[:, None] * p.out_pitch + torch.arange(Co)[None]
        if p.accumulate:
            out[oidx] = out[oidx] + acc
        else:
            out[oidx] = acc


def conv_batched(params_list):
    for p in params_list:
        conv(p)


def pack_weights_tc(w, phases, k, cin, cout, cout_pad, fmt=0, scale=1.0):
    return None


def pack_conv_weights(w, cin_pad, cout_pad, scale=1.0, sumsq=False):
    """e4s_pack_conv_weights_f32 restated with torch: [1, taps*cin_pad, cout_pad]."""
    co, ci, kh, kw = w.shape
    ws = w.float() * scale
    if sumsq:
        t = ws.pow(2).sum(dim=(2, 3)).t()[None]                                   # [1,Ci,Co]
        return torch.nn.functional.pad(t, (0, cout_pad - co, 0, cin_pad - ci)).contiguous()
    t = ws.permute(2, 3, 1, 0)                                                    # [kh,kw,Ci,Co]
    t = torch.nn.functional.pad(t, (0, cout_pad - co, 0, cin_pad - ci))
    return t.reshape(1, kh * kw * cin_pad, cout_pad).contiguous()


def pack_upconv_weights(w, fir, cout_pad, scale=1.0):
    """e4s_pack_upconv_weights_f32 restated with torch (the derivation of SURVEY.md appendix B.1): [4, 9*Ci, cout_pad]."""
    co, ci = w.shape[:2]
    ws = w.float() * scale
    kf = torch.flip(fir.float(), [0, 1])
    out = ws.new_zeros(2, 2, 3, 3, ci, co)
    for py in range(2):
        for u in range(3):
            for m in range(4):
                ky = 2 * (1 - u) + py + m - 1
                if not 0 <= ky <= 2:
                    continue
                for px in range(2):
                    for v in range(3):
                        for n in range(4):
                            kx = 2 * (1 - v) + px + n - 1
                            if 0 <= kx <= 2:
                                out[py, px, u, v] += ws[:, :, ky, kx].t() * kf[m, n]
    return torch.nn.functional.pad(out.reshape(4, 9 * ci, co), (0, cout_pad - co)).contiguous()


def upfirdn2d(x, kernel, up, down, pad0, pad1):
    from oracle import e4s_oracle as orc
    return orc.upfirdn2d(x, kernel, up, down, (pad0, pad1)).contiguous()


def upfirdn2d_general(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """e4s_upfirdn2d_f32 with its full parameter set: zero-insert, pad / crop per side, correlate with the flipped kernel, decimate."""
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    z = x.new_zeros(b, c, h * up_y, w * up_x)
    z[:, :, ::up_y, ::up_x] = x
    z = F.pad(z, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    z = z[:, :, max(-pad_y0, 0): z.shape[2] - max(-pad_y1, 0), max(-pad_x0, 0): z.shape[3] - max(-pad_x1, 0)]
    wt = torch.flip(kernel.to(x.dtype), [0, 1]).reshape(1, 1, kh, kw)
    hh, ww = z.shape[2], z.shape[3]
    y = F.conv2d(z.reshape(b * c, 1, hh, ww), wt).reshape(b, c, hh - kh + 1, ww - kw + 1)
    return y[:, :, ::down_y, ::down_x].contiguous()


def bias_act_grad(g, bias, ref, slope, scale):
    shape = [1, -1] + [1] * (g.ndim - 2)
    v = g if bias is None else g + bias.reshape(shape)
    return torch.where(ref > 0, v, v * slope) * scale


def bias_act(x, bias, slope, scale):
    shape = [1, -1] + [1] * (x.ndim - 2)
    v = x if bias is None else x + bias.reshape(shape)
    return torch.where(v < 0, v * slope, v) * scale


def noise_bias_act_nhwc(x, noise, noise_w, nsb, nsc, bias, slope, scale):
    b, h, w, c = x.shape
    v = x
    if noise is not None:
        nz = noise.reshape(noise.shape[0], noise.shape[1], h, w).permute(0, 2, 3, 1)
        v = v + noise_w * nz
    if bias is not None:
        v = v + bias
    x.copy_(torch.where(v < 0, v * slope, v) * scale)


def nchw_to_nhwc(x, c_pad=None):
    b, c, h, w = x.shape
    y = torch.zeros(b, h, w, c_pad or c)
    y[..., :c] = x.permute(0, 2, 3, 1)
    return y


def nhwc_to_nchw(x, c=None):
    c = c or x.shape[3]
    return x[..., :c].permute(0, 3, 1, 2).contiguous()


def mask_labels(mask, flags=None):
    nonzero = (mask != 0).sum(1)
    ones = (mask == 1).sum(1)
    bad = ~((nonzero == 1) & (ones == 1))
    labels = mask.argmax(1).to(torch.uint8)
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32)
    flags[0] = int(bad.any())
    return labels.contiguous(), flags


def labels_to_onehot(labels, k):
    return F.one_hot(labels.long(), k).permute(0, 3, 1, 2).float().contiguous()


def morphology(x, neighborhood, origin, border_value, dilate):
    """e4s_morphology_f32 restated with torch (window loops over the structuring element)."""
    se_h, se_w = neighborhood.shape
    h, w = x.shape[-2:]
    p = F.pad(x.float(), [origin[1], se_w - origin[1] - 1, origin[0], se_h - origin[0] - 1], mode="constant", value=border_value)
    out = None
    for i in range(se_h):
        for j in range(se_w):
            win = p[..., i:i + h, j:j + w]
            t = win + neighborhood[se_h - 1 - i, se_w - 1 - j] if dilate else win - neighborhood[i, j]
            out = t if out is None else (torch.maximum(out, t) if dilate else torch.minimum(out, t))
    return out.contiguous()


def tensor2im_u8(x, zero_center=True):
    v = x.float()
    if zero_center:
        v = (v + 1) / 2
    v = v.clamp(0, 1) * 255
    return v.permute(0, 2, 3, 1).to(torch.uint8).contiguous()


def im2tensor(x, want01=True, want_norm=True, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    x01 = x.permute(0, 3, 1, 2).float().div(255).contiguous()
    xn = ((x01 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)).contiguous()
    return (x01 if want01 else None), (xn if want_norm else None)


def depthwise_conv(x, weight, min_with_input=False):
    k = weight.shape[0]
    y = torch.nn.functional.conv2d(x, weight.view(1, 1, k, k).repeat(x.shape[1], 1, 1, 1), groups=x.shape[1], padding=k // 2)
    return torch.min(x, y) if min_with_input else y


def soft_erosion_finish(x, threshold):
    mask = x >= threshold
    x[mask] = 1.0
    x[~mask] /= x[~mask].max()
    return x, mask


def _per_plane(x, fn):
    from oracle import e4s_oracle as orc  # noqa: F401
    b, c = x.shape[:2]
    return torch.stack([torch.stack([torch.from_numpy(fn(x[i, j].numpy())) for j in range(c)]) for i in range(b)])


def pyr_down(x, round_u8=False):
    from oracle import e4s_oracle as orc
    return _per_plane(x, lambda a: orc.pyr_down(a, round_u8))


def pyr_up(x, other=None, mode=0):
    from oracle import e4s_oracle as orc
    up = _per_plane(x, orc.pyr_up)
    return up if mode == 0 else (other - up if mode == 1 else up + other)


def pyr_blend(la, lb, gm):
    return la * gm + lb * (1.0 - gm)


def swap_comp_styles(target, source, comp_mask, below_face):
    """e4s_swap_comp_styles_f32 restated with torch (mode per component: target / source / average)."""
    out = target.clone()
    k = target.shape[1]
    for c in range(k):
        mode = (comp_mask >> c) & 1
        if c == 7:
            mode = 2
        if c == 11:
            mode = 0
        if c == 8 and below_face:
            mode = 2
        if mode == 1:
            out[:, c] = source[:, c]
        elif mode == 2:
            out[:, c] = (target[:, c] + source[:, c]) / 2
    empty = source[:, 9].sum(dim=1) == 0
    out[empty, 9] = target[empty, 9]
    return out


def torgb(x_nhwc, cin, smod, wrgb, labels, regions, lab_hw, pixw, pixw_sb, bias, skip, fir, rgb, accumulate):
    b, h, w, _ = x_nhwc.shape
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    sy = _nearest(yy.reshape(-1), lab_hw[0], h) if lab_hw[0] else yy.reshape(-1) * 0
    sx = _nearest(xx.reshape(-1), lab_hw[1], w) if lab_hw[1] else xx.reshape(-1) * 0
    reg = labels.long()[:, sy, sx] if labels is not None else torch.zeros(b, h * w, dtype=torch.long)
    s = _f32(smod.data_ptr(), ((b - 1) * regions + int(reg.max()) + 1) * cin).reshape(-1, cin)
    rows = torch.arange(b)[:, None] * regions + reg
    xm = x_nhwc[..., :cin].reshape(b, h * w, cin) * s[rows]
    v = torch.einsum("bpc,oc->bop", xm, wrgb)
    if pixw is not None:
        m = _f32(pixw, (b - 1) * pixw_sb + lab_hw[0] * lab_hw[1])
        idx = torch.arange(b)[:, None] * pixw_sb + (sy * lab_hw[1] + sx)[None]
        v = v * m[idx][:, None, :]
    v = v.reshape(b, 3, h, w)
    if accumulate:
        rgb += v
        return
    if bias is not None:
        v = v + bias.reshape(1, 3, 1, 1)
    if skip is not None:
        from oracle import e4s_oracle as orc
        v = v + orc.upfirdn2d(skip, fir, up=2, pad=(2, 1))
    rgb.copy_(v)


def chan_stats(x_nhwc, c, eps=1e-5, want_rstd=True):
    v = x_nhwc[..., :c].double()
    mean = v.mean(dim=(1, 2))
    var = (v * v).mean(dim=(1, 2)) - mean * mean
    rstd = (1.0 / torch.sqrt(var.clamp(min=0) + eps)).float() if want_rstd else None
    return mean.float(), rstd


def vec_fc(x, w, scale, shift, act):
    y = x @ w.t()
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    return _act(y, act, 0, 1, None)


def residual_combine(a, c, a_stats=None, gate=None, gate_plus_one=False, r=None, r_sub=1, r_stats=None, relu=False,
                     prelu=None, out=None):
    b, h, w, _ = a.shape
    v = a[..., :c]
    if a_stats is not None:
        v = (v - a_stats[0][:, None, None]) * a_stats[1][:, None, None]
    if gate is not None:
        g = gate[:, None, None]
        v = v * g + v if gate_plus_one else v * g
    if r is not None:
        if r_sub > 1:
            q = r[:, ::r_sub, ::r_sub, :c][:, :h, :w]
        else:
            ys = _nearest(torch.arange(h), r.shape[1], h)
            xs = _nearest(torch.arange(w), r.shape[2], w)
            q = r[:, ys][:, :, xs][..., :c]
        if r_stats is not None:
            q = (q - r_stats[0][:, None, None]) * r_stats[1][:, None, None]
        v = v + q
    if relu:
        v = v.clamp(min=0)
    if prelu is not None:
        v = torch.where(v < 0, v * prelu, v)
    if out is None:
        out = torch.empty(b, h, w, c)
    out[..., :c] = v
    return out


def masked_mean(feat, c, mask, codes, c_off):
    b, h, w, _ = feat.shape
    ys = _nearest(torch.arange(h), mask.shape[2], h)
    xs = _nearest(torch.arange(w), mask.shape[3], w)
    m = (mask[:, :, ys][:, :, :, xs] != 0).double()
    tot = torch.einsum("bkhw,bhwc->bkc", m, feat[..., :c].double())
    area = m.sum(dim=(2, 3))
    codes[:, :, c_off:c_off + c] = torch.where(area[..., None] > 0, tot / area.clamp(min=1)[..., None], torch.zeros_like(tot)).float()


def mask_member_bits(mask):
    k = mask.shape[1]
    w = (2 ** torch.arange(k, dtype=torch.int64)).view(1, k, 1, 1)
    v = ((mask != 0).to(torch.int64) * w).sum(1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32)


def masked_mean_bits(feat, c, bits, k, codes, c_off):
    v = bits.to(torch.int64) & 0xFFFFFFFF
    mask = torch.stack([((v >> j) & 1).float() for j in range(k)], dim=1)
    masked_mean(feat, c, mask, codes, c_off)


def resize_bilinear_nchw_to_nhwc(x, hout, wout, c_pad, align_corners=False):
    y = F.interpolate(x, (hout, wout), mode="bilinear", align_corners=align_corners)
    return nchw_to_nhwc(y, c_pad)


def resize_bilinear_nhwc_to_nchw(x, c, hout, wout, align_corners=True):
    return F.interpolate(nhwc_to_nchw(x, c), (hout, wout), mode="bilinear", align_corners=align_corners)


def maxpool3x3s2(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def upsample_argmax(logits, c, hout, wout, lut=None):
    up = F.interpolate(nhwc_to_nchw(logits, c), (hout, wout), mode="bilinear", align_corners=True)
    lab = up.argmax(1).to(torch.uint8)
    return lut[lab.long()] if lut is not None else lab


def bicubic_down_norm(x, factor, taps, mean, std, c_pad, clamp=True):
    from oracle import e4s_oracle as orc
    y = orc.bicubic_downsample(x, factor)
    if clamp:
        y = y.clamp(0, 1)
    y = (y - mean.reshape(1, 3, 1, 1)) / std.reshape(1, 3, 1, 1)
    return nchw_to_nhwc(y, c_pad)


_NAMES = ["conv", "conv_batched", "pack_weights_tc", "pack_conv_weights", "pack_upconv_weights", "upfirdn2d", "upfirdn2d_general", "bias_act", "bias_act_grad", "noise_bias_act_nhwc", "nchw_to_nhwc", "nhwc_to_nchw",
          "mask_labels", "labels_to_onehot", "swap_comp_styles", "tensor2im_u8", "im2tensor", "morphology", "depthwise_conv", "soft_erosion_finish", "pyr_down", "pyr_up", "pyr_blend", "torgb", "chan_stats", "vec_fc", "residual_combine", "masked_mean", "mask_member_bits", "masked_mean_bits",
          "resize_bilinear_nchw_to_nhwc", "resize_bilinear_nhwc_to_nchw", "maxpool3x3s2", "upsample_argmax",
          "bicubic_down_norm"]


@contextmanager
def emulated():
    """Patch e4s2024_b200._lib's wrappers with the CPU emulation for the duration of a test."""
    saved = {n: getattr(L, n) for n in _NAMES}
    g = globals()
    for n in _NAMES:
        setattr(L, n, g[n])
    L.EMULATED = True                       # lets the public op functions accept CPU tensors for the duration of the test
    try:
        yield
    finally:
        L.EMULATED = False
        for n, f in saved.items():
            setattr(L, n, f)

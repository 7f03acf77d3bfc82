"""Differentiable (training-mode) path of the mask-guided generator: what `loss.backward()` through `net.G` needs for PTI fine-tuning
(reference training/video_swap_ft_coach.py:242-318, whose gradients flow through models/stylegan2/model.py:184-698
and op/conv2d_gradfix.py:134-225).

The forward pass of every convolution is the same fused kernel the inference path runs.  The backward pass follows the reference's
own formulation -- out = sum_k mask_k * demod_k * conv(x * s_k; W)  (model.py:395-398) differentiated region by region -- on this
library's kernels:
    data gradient    e4s_conv_tc / e4s_conv_f32 with the transposed (and, at the same resolution, flipped) weights; an up-convolution is
                     taken apart again into conv_transpose2d + Blur: blur^T = e4s_upfirdn2d_f32 with the flipped FIR, conv_transpose^T =
                     a stride-2 convolution
    weight gradient  e4s_conv_wgrad_f32
    style / demodulation table gradients   e4s_region_dot_f32;   mask multiply / demodulation of the incoming gradient  e4s_region_scale_f32
Small per-layer table algebra (bias adds, squares, rsqrt, the noise term) stays in torch autograd: it is O(batch x regions x channels).
Masks must be one-hot here (the pipelines' masks are); soft masks raise.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch.autograd import Function

from .. import _lib as L
from .. import engine as E
from ..engine import View
from .op import fused_leaky_relu, upfirdn2d


# Engine of the convolutions on this path: "tc" (tcgen05, operands split to 2^-16: fast, the default) or "f32" (exact-fp32 CUDA cores).
# Against fp64 autograd through the oracle, with the activation made linear, every gradient tensor of the generator agrees to 1.6e-4 of
# its largest entry on "tc" and 1.8e-5 on "f32"; with the real leaky ReLU both differ from fp64 by ~sqrt(P) where P is the fraction of
# pre-activations that rounding moves across the kink (tests/test_backward_gpu.py).  For scale: the reference's own convolutions run in
# TF32 (2^-11) on this GPU unless torch.backends.cudnn.allow_tf32 is switched off.
TRAIN_ENGINE = os.environ.get("E4S_TRAIN_ENGINE", "tc")


def set_train_engine(name: str):
    global TRAIN_ENGINE
    assert name in ("tc", "f32")
    TRAIN_ENGINE = name


def _r4(n: int) -> int:
    return (n + 3) // 4 * 4


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


class LinearFn(Function):
    """y[r, :] = (scale * W) x[r, :]  for x [R, In], W [Out, In]  (EqualLinear without its bias, model.py:156-162)."""

    @staticmethod
    def forward(ctx, x, weight, scale):
        x = x.contiguous().float()
        pw = E.pack_linear_weight(weight.detach().float(), scale=scale)
        y = E.linear_rows(x, x.shape[0], x.shape[1], 0, pw)
        ctx.save_for_backward(x, weight)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous().float()
        rows, n_in = x.shape
        n_out = weight.shape[0]
        dx = dw = None
        if ctx.needs_input_grad[0]:
            pwt = E.pack_linear_weight(weight.detach().float().t().contiguous(), scale=ctx.scale)
            dx = E.linear_rows(g, rows, n_out, 0, pwt)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(n_out, n_in, 1, 1, device=x.device, dtype=torch.float32)
            L.conv_wgrad(x.view(rows, 1, 1, n_in), n_in, g.view(rows, 1, 1, n_out), n_out, (1, 1), 1, 1, (1, 0, 1, 0, 0), None, ctx.scale, dw, False)
            dw = dw.view(n_out, n_in)
        return dx, dw, None


def equal_linear(layer, x: torch.Tensor) -> torch.Tensor:
    """Differentiable EqualLinear.forward (model.py:135-169) on [R, In] rows."""
    y = LinearFn.apply(x, layer.weight, layer.scale)
    if layer.activation:
        return fused_leaky_relu(y, layer.bias * layer.lr_mul)
    if layer.bias is not None:
        y = y + layer.bias * layer.lr_mul
    return y


class ModConvFn(Function):
    """y = d[b, r(p)] * conv(x * s[b, r(p)]; scale * W) for one-hot regions r(p) (ModulatedConv2d.forward, model.py:256-320, with the
    regional sum of StyledConv / ToRGB, :395-398 / :449-455); no noise / bias / activation.  NCHW in and out like the module."""

    @staticmethod
    def forward(ctx, x_nchw, weight, s, d, labels, mc, regions):
        x = L.nchw_to_nhwc(x_nchw.contiguous().float())
        conv, _ = mc.packed()
        out = E.conv(View(x), conv, up2=mc.upsample, smod=s.contiguous(), demod=None if d is None else d.contiguous(), regions=regions,
                     labels=labels, engine=TRAIN_ENGINE)
        ctx.mc, ctx.regions, ctx.labels = mc, regions, labels
        ctx.save_for_backward(x, weight, s, d if d is not None else torch.empty(0, device=x.device), out.t)
        return L.nhwc_to_nchw(out.t, conv.cout)

    @staticmethod
    def backward(ctx, gy):
        x, weight, s, d, y = ctx.saved_tensors
        mc, regions, labels = ctx.mc, ctx.regions, ctx.labels
        if d.numel() == 0:
            d = None
        b, hin, win, cin = x.shape
        cout, k, up = mc.out_channel, mc.kernel_size, mc.upsample
        ho, wo = y.shape[1], y.shape[2]
        g = L.nchw_to_nhwc(gy.contiguous().float(), _r4(cout))
        s = s.contiguous()
        # demodulation table: y = d * acc  =>  dL/dd[b,r,co] = sum_{p in r} g * acc = sum g * y / d
        dd = None
        if d is not None:
            dd = L.region_dot(g, y, cout, labels, regions) / d
        w0 = weight.detach()[0].float()                                            # [Co, Ci, k, k]
        if up:
            wt = w0.permute(1, 0, 2, 3).contiguous()                               # conv_transpose^T: stride-2 conv with the same taps
        else:
            wt = w0.permute(1, 0, 2, 3).flip(2, 3).contiguous()                    # conv^T: flipped taps
        pwt = E.pack_conv_weight(wt, cin_pad=_r8(cout), scale=mc.scale)
        # regions that occur in the mask (one host read per forward, cached on the label tensor by generator_forward)
        present = [0] if labels is None else (getattr(labels, "_e4s_present", None) or [int(v) for v in torch.unique(labels).tolist()])
        dx = torch.empty(b, hin, win, cin, device=x.device, dtype=torch.float32)
        dw = torch.empty(cout, cin, k, k, device=x.device, dtype=torch.float32)
        ds = torch.zeros(b, regions, cin, device=x.device, dtype=torch.float32)
        need_w = ctx.needs_input_grad[1]
        for n, r in enumerate(present):
            # incoming gradient of region r: mask multiply and demodulation (zero elsewhere), channels padded to the engines' cin % 8
            gr = L.region_scale(g, cout, d, labels, regions, select=r if labels is not None else -1, out_c=_r8(cout))
            if up:
                # Blur^T: upfirdn2d with the flipped FIR and pads (k-1-pad) = (2, 2): [2H, 2W] -> [2H+1, 2W+1]  (op/upfirdn2d.py:100-105)
                fk = torch.flip(mc.blur.kernel.detach().float(), [0, 1]).contiguous()
                gz = L.upfirdn2d(L.nhwc_to_nchw(gr, cout), fk, 1, 1, 2, 2)
                gsrc = L.nchw_to_nhwc(gz, _r8(cout))
                h = E.conv(View(gsrc), pwt, stride=2, pad=0, engine=TRAIN_ENGINE)
                geom, loop = (0, 0, 2, 1, 0), (hin, win)
            else:
                gsrc = gr
                h = E.conv(View(gsrc), pwt, stride=1, pad=k // 2, engine=TRAIN_ENGINE)
                geom, loop = (1, k // 2, 1, 0, 0), (ho, wo)
            sr = s[:, r]                                                           # [B, Ci] view, row stride regions * Ci
            ds[:, r] = L.region_dot(x, h.t, cin, None, 1)[:, 0]
            L.chan_scale_accum(h.t, cin, sr, dx, n > 0)
            if need_w:
                L.conv_wgrad(x, cin, gsrc, cout, loop, k, k, geom, sr, mc.scale, dw, n > 0)
        dxo = L.nhwc_to_nchw(dx, cin) if ctx.needs_input_grad[0] else None
        return dxo, (dw.view(1, cout, cin, k, k) if need_w else None), ds, dd, None, None, None


def tables(mc, style_rows: torch.Tensor, b: int, regions: int):
    """s [B,K,Ci] = modulation(style) (model.py:276), d [B,K,Co] = rsqrt(sum (scale W s)^2 + eps) (:279-281), differentiable."""
    s = equal_linear(mc.modulation, style_rows)                                    # [B*K, Ci]
    d = None
    if mc.demodulate:
        wsq = (mc.scale * mc.weight[0]).pow(2).sum((2, 3))                         # [Co, Ci]
        d = torch.rsqrt(LinearFn.apply(s * s, wsq, 1.0) + mc.eps).view(b, regions, mc.out_channel)
    return s.view(b, regions, mc.in_channel), d


def styled_conv(layer, x, style_rows, b, regions, labels, noise):
    """StyledConv.forward (model.py:351-423) for one-hot masks."""
    mc = layer.conv
    s, d = tables(mc, style_rows, b, regions)
    y = ModConvFn.apply(x, mc.weight, s, d, labels if (layer.mask_op and regions > 1) else None, mc, regions)
    if noise is None:
        noise = torch.empty(y.shape[0], 1, y.shape[2], y.shape[3], device=y.device, dtype=torch.float32).normal_()
    y = y + layer.noise.weight * noise
    return fused_leaky_relu(y, layer.activate.bias, layer.activate.negative_slope, layer.activate.scale)


def to_rgb(layer, x, style_rows, b, regions, labels, skip):
    """ToRGB.forward (model.py:426-479)."""
    mc = layer.conv
    s, _ = tables(mc, style_rows, b, regions)
    y = ModConvFn.apply(x, mc.weight, s, None, labels if (layer.mask_op and regions > 1) else None, mc, regions)
    y = y + layer.bias
    if skip is not None:
        y = y + upfirdn2d(skip, layer.upsample.kernel, up=layer.upsample.factor, down=1, pad=layer.upsample.pad)
    return y


def generator_forward(G, latent: torch.Tensor, mask: torch.Tensor, noise, structure_feats=None, use_structure_code=False):
    """Generator.forward (model.py:598-698) with input_is_latent=True, built from the differentiable pieces above.
    latent [B,K,n_latent,512]; returns (image, intermediate_feats)."""
    b, k, nl, sd = latent.shape
    ctx = E.RegionCtx(mask.to(latent.device), lazy=False)
    if not ctx.onehot:
        raise L.E4SError("the differentiable path needs one-hot masks (soft / overlapping masks are inference-only)")
    labels = ctx.labels
    labels._e4s_present = [int(v) for v in torch.unique(labels).tolist()]      # read once here instead of once per layer in the backward

    def regional(i):
        return latent[:, :, i].reshape(b * k, sd), k

    def glob(i):
        return latent[:, 0, i].reshape(b, sd), 1

    rl = G.remaining_layer_idx
    x = G.input.input.repeat(b, 1, 1, 1)
    st, kk = regional(0)
    out = styled_conv(G.conv1, x, st, b, kk, labels, noise[0])
    st, kk = regional(1)
    skip = to_rgb(G.to_rgb1, out, st, b, kk, labels, None)
    feats = None
    i = 1
    for conv1, conv2, n1, n2, rgb in zip(G.convs[::2], G.convs[1::2], noise[1::2], noise[2::2], G.to_rgbs):
        if i < rl:
            s1, s2 = regional(i), regional(i + 1)
            s3 = regional(i + 2) if (rl == 17 or i + 2 != rl) else glob(i + 2)
        else:
            s1, s2, s3 = glob(i), glob(i + 1), glob(i + 2)
        out = styled_conv(conv1, out, s1[0], b, s1[1], labels, n1)
        if i < rl and i + 2 == G.split_layer_idx:
            if use_structure_code:
                out = structure_feats
            feats = out
        out = styled_conv(conv2, out, s2[0], b, s2[1], labels, n2)
        skip = to_rgb(rgb, out, s3[0], b, s3[1], labels, skip)
        i += 2
    return skip, feats

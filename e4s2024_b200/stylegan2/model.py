"""Mask-guided StyleGAN2 generator -- drop-in for the reference `models/stylegan2/model.py`
(Generator side: :14-53, :78-94, :135-169, :184-698).  Same class names, constructor
signatures, forward() signatures, return tuples and state_dict keys; the computation is the
B200 engine (fused gather -> modulate -> GEMM -> demod/noise/bias/act epilogue), NHWC inside.

What changed structurally versus the reference (results identical within fp32 tolerance):
  * one fused convolution per layer instead of K = 12 grouped convolutions + mask-multiply-adds:
    output pixel p is computed once with the style of ITS region r(p)  (exact for one-hot masks;
    soft / overlapping masks take a per-region accumulate path on the same kernels);
  * conv_transpose2d(stride 2) + Blur is folded into four 3x3 phase filters (engine.polyphase_weights);
  * per-sample weights are never materialised: modulation scales the gathered activations,
    demodulation is a per-(sample, region, channel) epilogue scale;
  * noise, bias, leaky-ReLU*sqrt(2) run in the conv epilogue; ToRGB fuses bias and the FIR-upsampled skip.
Forward only (inference hot path); Discriminator / ConvLayer / ResBlock are training-only and not provided.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..engine import View
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

SQRT2 = 2 ** 0.5


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """model.py:34-53."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Blur(nn.Module):
    """model.py:78-94."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


def _ver(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


class EqualLinear(nn.Module):
    """model.py:135-169."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul
        self._pack = None

    def packed(self):
        key = _ver(self.weight) + (_ver(self.bias) if self.bias is not None else ())
        if self._pack is None or self._pack[0] != key:
            pw = E.pack_linear_weight(self.weight.detach(), scale=self.scale)
            b = None if self.bias is None else (self.bias.detach() * self.lr_mul).contiguous()
            self._pack = (key, pw, b)
        return self._pack[1], self._pack[2]

    def rows(self, src: torch.Tensor, rows: int, row_stride: int, offset: int) -> torch.Tensor:
        pw, b = self.packed()
        if self.activation:
            return E.linear_rows(src, rows, row_stride, offset, pw, bias=b, act=L.ACT_LRELU, slope=0.2, gain=SQRT2)
        return E.linear_rows(src, rows, row_stride, offset, pw, bias=b)

    def forward(self, input):
        x = input.contiguous().float()
        lead = x.shape[:-1]
        y = self.rows(x, x.numel() // x.shape[-1], x.shape[-1], 0)
        return y.reshape(*lead, -1)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class StyleRows:
    """`rows` style vectors of length 512 inside a contiguous tensor: row r starts at offset + r*stride."""
    __slots__ = ("t", "rows", "stride", "offset", "regions")

    def __init__(self, t, rows, stride, offset, regions):
        self.t, self.rows, self.stride, self.offset, self.regions = t, rows, stride, offset, regions

    @staticmethod
    def from_tensor(style: torch.Tensor) -> "StyleRows":
        s = style.contiguous().float()
        if s.dim() == 2:
            return StyleRows(s, s.shape[0], s.shape[1], 0, 1)
        assert s.dim() == 3
        return StyleRows(s, s.shape[0] * s.shape[1], s.shape[2], 0, s.shape[1])


class ModulatedConv2d(nn.Module):
    """model.py:184-320."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1], fused=True):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        fan_in = in_channel * kernel_size ** 2
        self.scale = 1 / math.sqrt(fan_in)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self.fused = fused
        self._pack = None

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    # ---- packed weights (rebuilt when the parameters change, e.g. after PTI fine-tuning) ----------
    def packed(self):
        key = _ver(self.weight) + ((_ver(self.blur.kernel)) if self.upsample else ())
        if self._pack is None or self._pack[0] != key:
            w = self.weight.detach()[0].float()                                   # [Co,Ci,k,k]; scale is applied by the pack kernels
            if self.upsample:
                if self.kernel_size != 3 or tuple(self.blur.kernel.shape) != (4, 4) or self.blur.pad != (1, 1):
                    raise L.E4SError("up-sampling ModulatedConv2d supports kernel_size=3 with a 4-tap blur")
                conv = E.pack_up_weight(w, self.blur.kernel, scale=self.scale)
            else:
                conv = E.pack_conv_weight(w, scale=self.scale)
            wsq = None
            if self.demodulate:                                                   # sum_taps (scale*W)^2 -> [Ci] x [Co]
                wsq = E.pack_conv_weight(w, scale=self.scale, sumsq=True)
            self._pack = (key, conv, wsq)
        return self._pack[1], self._pack[2]

    def tables(self, st: StyleRows):
        """s[row, ci] = modulation(style) (model.py:276); d[row, co] = rsqrt(sum (scale*W*s)^2 + eps) (:279-281)."""
        _, wsq = self.packed()
        s = self.modulation.rows(st.t, st.rows, st.stride, st.offset)
        d = None
        if self.demodulate:
            d = E.linear_rows(s, st.rows, self.in_channel, 0, wsq, act=L.ACT_RSQRT_EPS, slope=self.eps, in_square=True)
        return s, d

    def run(self, x: View, st: StyleRows, ctx: Optional[E.RegionCtx], tables=None, **epilogue) -> View:
        """Fused modulated convolution on an NHWC view; `epilogue` = noise/bias/activation kwargs of engine.conv.
        `tables` = precomputed (s, d) (Generator.forward batches all layers' table GEMMs into two launches)."""
        if self.downsample:
            raise NotImplementedError("downsample=True is only used by the (training-only) Discriminator")
        conv, _ = self.packed()
        s, d = tables if tables is not None else self.tables(st)
        regional = st.regions > 1
        if regional and ctx is None:
            raise L.E4SError("regional styles need a mask")
        if not regional or ctx.onehot:
            rj = uz = None
            if regional:
                _, h, w = x.bhw
                ho, wo = (2 * h, 2 * w) if self.upsample else (h, w)
                if self.upsample and conv.tc9 is not None:
                    uz = ctx.upz_for(h, w)          # conv_transpose as a (cell, region) GEMM + FIR pass (csrc/conv_tc_upz.cu)
                if uz is None:
                    rj = ctx.jobs_for(ho, wo, self.upsample, E.wide_eligible(self.in_channel, self.out_channel, h, w, self.upsample))
            return E.conv(x, conv, up2=self.upsample, smod=s, demod=d, regions=st.regions,
                          labels=ctx.labels if regional else None, region_jobs=rj, upz=uz, **epilogue)
        # generic float masks: sum_k mask_k * conv(x; style_k), then the epilogue (model.py:395-398)
        out = None
        for r in range(st.regions):
            out = E.conv(x, conv, up2=self.upsample, smod=s, demod=d, regions=st.regions, smod_off=r * self.in_channel,
                         demod_off=r * self.out_channel, pixw=(ctx.mask, r), out=out, accumulate=r > 0)
        if epilogue:
            b, h, w = out.bhw
            noise, nw = epilogue.get("noise"), epilogue.get("noise_w")
            nsb = nsc = 0
            if noise is not None:
                nsb = 0 if noise.shape[0] == 1 else noise.shape[1] * h * w
                nsc = 0 if noise.shape[1] == 1 else h * w
            L.noise_bias_act_nhwc(out.t, noise, nw, nsb, nsc, epilogue.get("ch_shift"), epilogue.get("slope", 1.0),
                                  epilogue.get("gain", 1.0))
        return out

    def forward(self, input, style):
        x = View(L.nchw_to_nhwc(input.contiguous().float()))
        out = self.run(x, StyleRows.from_tensor(style), None)
        return L.nhwc_to_nchw(out.t)


class NoiseInjection(nn.Module):
    """model.py:323-335 (standalone form; inside StyledConv the add happens in the conv epilogue)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """model.py:338-348."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


def _noise_for(noise, b, h, w, device):
    if noise is None:                                   # model.py:329-333: fresh N(0,1) per call
        return torch.empty(b, 1, h, w, device=device, dtype=torch.float32).normal_()
    return noise.detach().to(device).contiguous().float()


class StyledConv(nn.Module):
    """model.py:351-423."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True, mask_op=False):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)
        self.mask_op = mask_op

    def run(self, x: View, st: StyleRows, ctx, noise, tables=None, **extra) -> View:
        b, h, w = x.bhw
        if self.conv.upsample:
            h, w = 2 * h, 2 * w
        nz = _noise_for(noise, b, h, w, x.t.device)
        return self.conv.run(x, st, ctx if self.mask_op else None, tables=tables, noise=nz, noise_w=self.noise.weight.detach(),
                             ch_shift=self.activate.bias.detach(), act=L.ACT_LRELU,
                             slope=self.activate.negative_slope, gain=self.activate.scale, **extra)

    def forward(self, input, style, mask, noise=None):
        x = View(L.nchw_to_nhwc(input.contiguous().float()))
        ctx = None
        if self.mask_op:
            mc = self.conv
            uk = [(x.bhw[1], x.bhw[2])] if (mc.upsample and E.upz_eligible(mc.in_channel, mc.out_channel)) else []
            ctx = E.RegionCtx(mask, upz_keys=uk)
        return L.nhwc_to_nchw(self.run(x, StyleRows.from_tensor(style), ctx, noise).t)


class ToRGB(nn.Module):
    """model.py:426-479."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1], mask_op=False):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))
        self.mask_op = mask_op
        self._wrgb = None

    def _weights(self):
        key = _ver(self.conv.weight)
        if self._wrgb is None or self._wrgb[0] != key:
            w = (self.conv.weight.detach()[0, :, :, 0, 0] * self.conv.scale).contiguous().float()    # [3,Ci]
            self._wrgb = (key, w)
        return self._wrgb[1]

    def fused_args(self, b: int, h: int, w: int, s: torch.Tensor, skip: Optional[torch.Tensor], device) -> dict:
        """Arguments of the ToRGB tail fused into the preceding StyledConv's epilogue (engine.conv `rgb=`): this layer's
        1x1 modulated conv + bias + FIR-upsampled skip (model.py:439-479) computed from the activations in registers."""
        fir = None
        if skip is not None:
            fir = self.upsample.kernel.detach().float().contiguous()
            if tuple(fir.shape) != (4, 4) or self.upsample.pad != (2, 1):
                raise L.E4SError("ToRGB skip path supports the 4-tap FIR (pad=(2,1)) only")
            skip = skip.contiguous().float()
        return {"rgb": torch.empty(b, 3, h, w, device=device, dtype=torch.float32), "w": self._weights(), "smod": s,
                "bias": self.bias.detach().reshape(3).contiguous(), "skip": skip, "fir": fir}

    def run(self, x: View, st: StyleRows, ctx, skip: Optional[torch.Tensor], tables=None) -> torch.Tensor:
        """x NHWC view, skip NCHW [B,3,H/2,W/2] or None -> rgb NCHW [B,3,H,W]."""
        b, h, w = x.bhw
        s = tables[0] if tables is not None else self.conv.modulation.rows(st.t, st.rows, st.stride, st.offset)
        rgb = torch.empty(b, 3, h, w, device=x.t.device, dtype=torch.float32)
        fir = None
        if skip is not None:
            fir = self.upsample.kernel
            if tuple(fir.shape) != (4, 4) or self.upsample.pad != (2, 1):
                raise L.E4SError("ToRGB skip path supports the 4-tap FIR (pad=(2,1)) only")
            skip = skip.contiguous().float()
            assert skip.shape == (b, 3, h // 2, w // 2), skip.shape
        bias = self.bias.detach().reshape(3).contiguous()
        regional = self.mask_op and st.regions > 1
        xin = View(x.t, x.c, x.coff)
        if not regional:
            L.torgb(xin.t, x.c, s, self._weights(), None, st.regions, (0, 0), None, 0, bias, skip, fir, rgb, False)
        elif ctx.onehot:
            L.torgb(xin.t, x.c, s, self._weights(), ctx.labels, st.regions, ctx.lab_hw, None, 0, bias, skip, fir, rgb, False)
        else:
            m = ctx.mask
            plane = m.shape[2] * m.shape[3]
            for r in range(st.regions):
                L.torgb(xin.t, x.c, s[r:], self._weights(), None, st.regions, ctx.lab_hw, m.data_ptr() + 4 * r * plane,
                        m.shape[1] * plane, bias, skip, fir, rgb, r > 0)
        return rgb

    def forward(self, input, style, mask, skip=None):
        x = View(L.nchw_to_nhwc(input.contiguous().float()))
        st = StyleRows.from_tensor(style)
        ctx = E.RegionCtx(mask) if (self.mask_op and st.regions > 1) else None
        return self.run(x, st, ctx, skip)


class Generator(nn.Module):
    """model.py:482-698."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 split_layer_idx=7, remaining_layer_idx=18):
        super().__init__()
        self.split_layer_idx = split_layer_idx
        self.remaining_layer_idx = remaining_layer_idx
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
                         256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel, mask_op=True)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False, mask_op=True)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        rl = self.remaining_layer_idx
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            conv_masked = not (i > (2 + rl // 2))
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel,
                                         mask_op=conv_masked))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel,
                                         mask_op=conv_masked))
            self.to_rgbs.append(ToRGB(out_channel, style_dim, mask_op=not (rl != 17 and i >= (2 + rl // 2))))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        E.install_pack_invalidation(self)
        # The drop-in is constructed in eval() mode (every pipeline of the reference calls .eval() on the loaded network anyway): train()
        # is the explicit switch to the differentiable path, so a caller that merely forgot eval() / no_grad() does not pay for autograd.
        self.eval()

    def _region_job_keys(self):
        """(hout, wout, up2) of the masked StyledConv layers whose geometry the halo kernel takes."""
        keys = []
        res = 4
        for conv_up, conv2 in zip(self.convs[::2], self.convs[1::2]):
            res *= 2
            if conv_up.mask_op and not E.upz_eligible(conv_up.conv.in_channel, conv_up.conv.out_channel, True) and \
                    E.halo_geometry_ok(conv_up.conv.in_channel, conv_up.conv.out_channel, res // 2, res // 2, True):
                keys.append((res, res, True))
            if conv2.mask_op and E.halo_geometry_ok(conv2.conv.in_channel, conv2.conv.out_channel, res, res, False):
                keys.append((res, res, False))
        return keys

    def _upz_keys(self):
        """(hin, win) of the masked up-convolutions that run as the (cell, region) conv_transpose GEMM."""
        keys = []
        res = 4
        for conv_up in self.convs[::2]:
            if conv_up.mask_op and E.upz_eligible(conv_up.conv.in_channel, conv_up.conv.out_channel):
                keys.append((res, res))
            res *= 2
        return keys

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def forward(self, styles, structure_feats, mask, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True,
                use_structure_code=False, _ctx=None, _host_flag=None):
        anchor = grad_anchor(self, (styles,))
        if anchor is not None and self.training and anchor.is_cuda and _ctx is None:
            # train() mode + autograd recording (the PTI coach: training/video_swap_ft_coach.py:242-318): the differentiable path
            return self._forward_train(styles, structure_feats, mask, return_latents, truncation, input_is_latent, noise, randomize_noise,
                                       use_structure_code)
        with torch.no_grad():
            out = self._forward(styles, structure_feats, mask, return_latents, inject_index, truncation, truncation_latent,
                                input_is_latent, noise, randomize_noise, use_structure_code, _ctx, _host_flag)
        return inference_only(out, anchor)

    def _forward_train(self, styles, structure_feats, mask, return_latents, truncation, input_is_latent, noise, randomize_noise,
                       use_structure_code):
        from . import grad as GR
        if not input_is_latent or truncation < 1 or len(styles) != 1 or styles[0].ndim != 4:
            raise L.E4SError("the differentiable path takes styles=[latent [B,K,n_latent,512]] with input_is_latent=True (what Net3.gen_img passes)")
        latent = styles[0].float()
        if latent.shape[3] != self.style_dim or latent.shape[2] < self.n_latent:
            raise L.E4SError(f"latent shape {tuple(latent.shape)} incompatible with n_latent={self.n_latent}")
        if noise is None:
            noise = [None] * self.num_layers if randomize_noise else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        E.invalidate_packs(self)          # optimisers update weights in place through .data (no version bump): never trust a cached packing here
        image, feats = GR.generator_forward(self, latent, mask, noise, structure_feats, use_structure_code)
        return (image, latent, feats) if return_latents else (image, None, feats)

    def _forward(self, styles, structure_feats, mask, return_latents, inject_index, truncation, truncation_latent, input_is_latent,
                 noise, randomize_noise, use_structure_code, _ctx, _host_flag):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            noise = [None] * self.num_layers if randomize_noise else \
                [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) != 1 or styles[0].ndim != 4:
            # model.py:642-659: the 2-/3-D and style-mixing branches index latent[:, :, i] and cannot run with
            # regional styles in the reference either; only the [B,K,n_latent,512] form is meaningful.
            raise L.E4SError("Generator expects styles=[latent] with latent of shape [B, K, n_latent, style_dim]")
        latent = styles[0].contiguous().float()
        b, k, nl, sd = latent.shape
        if sd != self.style_dim or nl < self.n_latent:
            raise L.E4SError(f"latent shape {tuple(latent.shape)} incompatible with n_latent={self.n_latent}")
        # lazy region context: no device->host read while the layers are being enqueued (the one-hot assumption is checked
        # after the last launch; a soft / overlapping mask re-runs the forward on the generic per-region path)
        if _host_flag is None:
            _host_flag = getattr(self, "_capture_host_flag", None)      # set by serving.GraphedSwapPath while it captures / replays
        ctx = _ctx if _ctx is not None else E.RegionCtx(mask.to(latent.device), self._region_job_keys(), lazy=True, host_flag=_host_flag,
                                                              upz_keys=self._upz_keys())
        if ctx.k != k or ctx.mask.shape[0] != b:
            raise L.E4SError("mask and latent disagree on batch / number of regions")

        def regional(i):      # latent[:, :, i]
            return StyleRows(latent, b * k, nl * sd, i * sd, k)

        def glob(i):          # latent[:, 0, i]
            return StyleRows(latent, b, k * nl * sd, i * sd, 1)

        # which style rows feed which layer (model.py:661-690), then ALL style tables in two batched launches
        rl = self.remaining_layer_idx
        plan = [(self.conv1.conv, regional(0)), (self.to_rgb1.conv, regional(1))]
        i = 1
        for conv1, conv2, to_rgb in zip(self.convs[::2], self.convs[1::2], self.to_rgbs):
            if i < rl:
                st3 = regional(i + 2) if (rl == 17 or i + 2 != rl) else glob(i + 2)
                plan += [(conv1.conv, regional(i)), (conv2.conv, regional(i + 1)), (to_rgb.conv, st3)]
            else:
                plan += [(conv1.conv, glob(i)), (conv2.conv, glob(i + 1)), (to_rgb.conv, glob(i + 2))]
            i += 2
        tabs = _batched_tables(plan)
        sty = {id(mc): st for mc, st in plan}

        def sconv(layer, xin, nz):
            return layer.run(xin, sty[id(layer.conv)], ctx, nz, tables=tabs[id(layer.conv)])

        def rgb(layer, xin, skip_):
            return layer.run(xin, sty[id(layer.conv)], ctx, skip_, tables=tabs[id(layer.conv)])

        x = View(L.nchw_to_nhwc(self.input.input.detach().float().repeat(b, 1, 1, 1).contiguous()))
        out = sconv(self.conv1, x, noise[0])
        skip = rgb(self.to_rgb1, out, None)
        intermediate_feats = None
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                        self.to_rgbs):
            out = sconv(conv1, out, noise1)
            if i < rl and i + 2 == self.split_layer_idx:
                if use_structure_code:
                    out = View(L.nchw_to_nhwc(structure_feats.contiguous().float()))
                intermediate_feats = L.nhwc_to_nchw(out.t)
            st2, st3 = sty[id(conv2.conv)], sty[id(to_rgb.conv)]
            _, hh, ww = out.bhw
            if st2.regions == 1 and st3.regions == 1 and skip is not None and \
                    E.rgb_fusable(conv2.conv.in_channel, conv2.conv.out_channel, hh, ww):
                # un-masked resolutions: ToRGB rides in the conv epilogue (the feature map is not read again; the last
                # layer's 1024^2 x 32 activations are never written at all)
                fa = to_rgb.fused_args(b, hh, ww, tabs[id(to_rgb.conv)][0], skip, out.t.device)
                last = to_rgb is self.to_rgbs[-1]
                out = conv2.run(out, st2, ctx, noise2, tables=tabs[id(conv2.conv)], rgb=fa, store_out=not last)
                skip = fa["rgb"]
            else:
                out = sconv(conv2, out, noise2)
                skip = rgb(to_rgb, out, skip)
            i += 2
        image = skip
        if not ctx.verify():
            # the mask was not one-hot: everything above used the label fast path -> redo with a synchronous context
            # (which knows) on the generic float-mask path; `noise` / styles are already resolved, so the rerun sees the same inputs
            sync_ctx = E.RegionCtx(mask.to(latent.device), self._region_job_keys(), lazy=False, upz_keys=self._upz_keys())
            return self.forward([latent], structure_feats, mask, return_latents=return_latents, input_is_latent=True, noise=noise,
                                randomize_noise=randomize_noise, use_structure_code=use_structure_code, _ctx=sync_ctx)
        if return_latents:
            return image, latent, intermediate_feats
        return image, None, intermediate_feats


_NO_BACKWARD = ("this forward ran on the inference-only path: gradients flow through Generator / Net3.gen_img / Net3.cal_style_codes only in "
                "train() mode on CUDA tensors (as the PTI coach uses them); the encoder and the face parser have no backward.")


class _InferenceOnly(torch.autograd.Function):
    """Identity whose backward explains itself.  The kernels run under no_grad, so a caller that expects gradients (the reference video
    pipeline's PTI step, training/video_swap_ft_coach.py:242-318: `loss.backward()` through net.G) would otherwise die inside autograd
    with 'element 0 of tensors does not require grad'; forward-only callers are unaffected."""

    @staticmethod
    def forward(ctx, out, anchor):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, grad):
        raise L.E4SError(_NO_BACKWARD)


def grad_anchor(module: nn.Module, inputs=()):
    """A tensor that requires grad among the inputs / parameters when autograd is recording, else None."""
    if not torch.is_grad_enabled():
        return None
    for t in inputs:
        for u in (t if isinstance(t, (list, tuple)) else (t,)):
            if isinstance(u, torch.Tensor) and u.requires_grad:
                return u
    for p_ in module.parameters():
        if p_.requires_grad:
            return p_
    return None


def inference_only(outputs, anchor):
    """Attach the explanatory backward to every floating-point tensor of `outputs` (a tuple as the modules return it)."""
    if anchor is None:
        return outputs
    if isinstance(outputs, torch.Tensor):
        return _InferenceOnly.apply(outputs, anchor)
    return tuple(_InferenceOnly.apply(o, anchor) if isinstance(o, torch.Tensor) and o.is_floating_point() else o for o in outputs)


def _batched_tables(plan):
    """plan: [(ModulatedConv2d, StyleRows)] -> {id(module): (s, d)} with all modulation GEMMs in one batched launch
    and all demodulation GEMMs in a second one (they were 43 tiny launches per forward)."""
    ps, ss = [], []
    for mc, st in plan:
        pw, bias = mc.modulation.packed()
        s_, p_ = E.linear_rows(st.t, st.rows, st.stride, st.offset, pw, bias=bias, launch=False)
        ps.append(p_)
        ss.append(s_)
    L.conv_batched(ps)
    pd, tabs = [], {}
    for (mc, st), s_ in zip(plan, ss):
        d_ = None
        if mc.demodulate:
            _, wsq = mc.packed()
            d_, p_ = E.linear_rows(s_, st.rows, mc.in_channel, 0, wsq, act=L.ACT_RSQRT_EPS, slope=mc.eps, in_square=True,
                                   launch=False)
            pd.append(p_)
        tabs[id(mc)] = (s_, d_)
    if pd:
        L.conv_batched(pd)
    return tabs


def generator_state_shapes(size: int, style_dim: int = 512, n_mlp: int = 8, **kw):
    """{name: shape} of Generator.state_dict() without allocating it (tests / synthetic weights)."""
    with torch.device("meta"):
        g = Generator(size, style_dim, n_mlp, **kw)
    return {k: tuple(v.shape) for k, v in g.state_dict().items()}

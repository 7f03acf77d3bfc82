"""CPU oracle for the E4S hot path -- TEST INFRASTRUCTURE ONLY.

A functional, state-dict driven restatement (plain PyTorch CPU ops, fp32 or fp64)
of the reference algorithms on the hot path.  Nothing in the product package
(`e4s2024_b200/`) may import this file; only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s baseline legs do (cpu_baseline, `--impl reference`, and `gpu_reference`:
the same torch ops -- the reference's K grouped cuDNN convolutions per masked layer --
executed on the B200 as the "existing Blackwell path" the new kernels are compared with;
every function follows the device of its inputs).

Parity pinning: the reference holds no golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against the reference's own modules executed in
the build container (`oracle/make_golden.py` imports /root/reference with CPU shims
and writes `tests/golden/*.npz`); `tests/test_oracle_golden.py` replays them.

Every function cites the reference file:line it restates (paths relative to the
reference root).  Weights are taken from a state dict that uses the reference's
own parameter names, so a reference checkpoint loads unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# small ops (models/stylegan2/op, models/stylegan2/model.py:23-53,78-94,135-169)
# --------------------------------------------------------------------------------------


def fir_kernel(taps: Sequence[float], gain: float = 1.0, dtype=torch.float32, device=None) -> torch.Tensor:
    """models/stylegan2/model.py:23-31 (make_kernel): outer product, normalised to sum 1."""
    k = torch.tensor(list(taps), dtype=dtype, device=device)
    k2 = torch.outer(k, k)
    return k2 / k2.sum() * gain


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1,
              pad: Tuple[int, int] = (0, 0)) -> torch.Tensor:
    """Zero-insert upsample, pad (negative = crop), correlate with the FLIPPED kernel, decimate.

    Semantics of models/stylegan2/op/upfirdn2d_kernel.cu:71-134 (flip at :77) and the
    pure-torch text in swap_face_fine/ops/upfirdn2d/upfirdn2d.py:163-193; same (pad0,pad1)
    on both axes as models/stylegan2/op/upfirdn2d.py:142-147.
    """
    b, c, h, w = x.shape
    p0, p1 = pad
    kh, kw = kernel.shape
    z = x.new_zeros(b, c, h * up, w * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    z = z[:, :, max(-p0, 0): z.shape[2] - max(-p1, 0), max(-p0, 0): z.shape[3] - max(-p1, 0)]
    wt = torch.flip(kernel.to(x.dtype), [0, 1]).reshape(1, 1, kh, kw)
    hh, ww = z.shape[2], z.shape[3]
    y = F.conv2d(z.reshape(b * c, 1, hh, ww), wt)
    y = y.reshape(b, c, hh - kh + 1, ww - kw + 1)
    return y[:, :, ::down, ::down]


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                     scale: float = 2 ** 0.5) -> torch.Tensor:
    """models/stylegan2/op/fused_bias_act_kernel.cu:26-47 case 30: lrelu(x + b[c]) * scale."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.reshape(shape).to(x.dtype), negative_slope) * scale


def equal_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                 lr_mul: float = 1.0, activation: bool = False) -> torch.Tensor:
    """models/stylegan2/model.py:135-164 (EqualLinear.forward)."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    w = weight.to(x.dtype) * scale
    if activation:
        return fused_leaky_relu(F.linear(x, w), bias.to(x.dtype) * lr_mul)
    return F.linear(x, w, None if bias is None else bias.to(x.dtype) * lr_mul)


def nearest_resize(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """F.interpolate(mode='nearest') = legacy floor(dst*in/out) (model.py:391,445)."""
    return F.interpolate(x, size=size, mode="nearest")


# --------------------------------------------------------------------------------------
# StyleGAN2 regional generator (models/stylegan2/model.py:184-698)
# --------------------------------------------------------------------------------------


def modulated_conv2d(x: torch.Tensor, style: torch.Tensor, sd: SD, prefix: str, *, demodulate: bool = True,
                     upsample: bool = False, blur_taps: Sequence[float] = (1, 3, 3, 1)) -> torch.Tensor:
    """models/stylegan2/model.py:242-320 (fused branch :276-318).

    `prefix` names a ModulatedConv2d, e.g. 'convs.3.conv.'.
    Per-sample weights W*s (+demod, eps 1e-8 inside rsqrt) -> grouped conv; the up
    variant is conv_transpose2d(stride 2) followed by Blur(pad=(1,1), gain 4).
    """
    weight = sd[prefix + "weight"].to(x.dtype)            # [1, Co, Ci, k, k]
    _, co, ci, k, _ = weight.shape
    b, _, h, w = x.shape
    s = equal_linear(style, sd[prefix + "modulation.weight"], sd[prefix + "modulation.bias"])
    wmod = (1.0 / math.sqrt(ci * k * k)) * weight * s.reshape(b, 1, ci, 1, 1)
    if demodulate:
        wmod = wmod * torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8).reshape(b, co, 1, 1, 1)
    if upsample:
        wt = wmod.transpose(1, 2).reshape(b * ci, co, k, k)
        y = F.conv_transpose2d(x.reshape(1, b * ci, h, w), wt, padding=0, stride=2, groups=b)
        y = y.reshape(b, co, y.shape[2], y.shape[3])
        p = (len(blur_taps) - 2) - (k - 1)
        return upfirdn2d(y, fir_kernel(blur_taps, 4.0, x.dtype, x.device), pad=((p + 1) // 2 + 1, p // 2 + 1))
    y = F.conv2d(x.reshape(1, b * ci, h, w), wmod.reshape(b * co, ci, k, k), padding=k // 2, groups=b)
    return y.reshape(b, co, y.shape[2], y.shape[3])


def _regional(x, style, mask, fn, out_hw):
    """Sum over regions of mask_k * f(x; style_k)  (model.py:391-398, :445-454)."""
    seg = nearest_resize(mask.to(x.dtype), out_hw)
    acc = None
    for k in range(style.shape[1]):
        yk = fn(x, style[:, k]) * seg[:, k:k + 1]
        acc = yk if acc is None else acc + yk
    return acc


def styled_conv(x, style, mask, sd: SD, prefix: str, *, upsample: bool, mask_op: bool,
                noise: Optional[torch.Tensor]) -> torch.Tensor:
    """models/stylegan2/model.py:382-423 (StyledConv.forward). `noise` must be given (pinned)."""
    fn = lambda xx, st: modulated_conv2d(xx, st, sd, prefix + "conv.", upsample=upsample)
    if mask_op:
        h, w = x.shape[2:]
        y = _regional(x, style, mask, fn, (2 * h, 2 * w) if upsample else (h, w))
    else:
        y = fn(x, style)
    y = y + sd[prefix + "noise.weight"].to(x.dtype) * noise.to(x.dtype)
    return fused_leaky_relu(y, sd[prefix + "activate.bias"])


def to_rgb(x, style, mask, skip, sd: SD, prefix: str, *, mask_op: bool,
           blur_taps: Sequence[float] = (1, 3, 3, 1)) -> torch.Tensor:
    """models/stylegan2/model.py:439-479 (ToRGB.forward); skip is upsampled with
    Upsample(:34-53): upfirdn2d(up=2, kernel*4, pad=(2,1))."""
    fn = lambda xx, st: modulated_conv2d(xx, st, sd, prefix + "conv.", demodulate=False)
    y = _regional(x, style, mask, fn, tuple(x.shape[2:])) if mask_op else fn(x, style)
    y = y + sd[prefix + "bias"].to(x.dtype)
    if skip is not None:
        y = y + upfirdn2d(skip, fir_kernel(blur_taps, 4.0, x.dtype, x.device), up=2, pad=(2, 1))
    return y


def generator_channels(channel_multiplier: int = 2) -> Dict[int, int]:
    """models/stylegan2/model.py:512-522."""
    m = channel_multiplier
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m, 1024: 16 * m}


def generator_forward(sd: SD, size: int, latent: torch.Tensor, mask: torch.Tensor,
                      noise: Optional[List[torch.Tensor]] = None, *, split_layer_idx: int = 7,
                      remaining_layer_idx: int = 18, prefix: str = "") -> Tuple[torch.Tensor, torch.Tensor]:
    """models/stylegan2/model.py:607-698 with input_is_latent=True, a 4-D latent [B,K,n_latent,512],
    randomize_noise=False (noise=None -> registered buffers noises.noise_i) or an explicit list.

    Returns (image [B,3,size,size], intermediate_feats).
    """
    log_size = int(math.log2(size))
    num_layers = (log_size - 2) * 2 + 1
    rl = remaining_layer_idx
    dt = latent.dtype
    if noise is None:
        noise = [sd[f"{prefix}noises.noise_{i}"] for i in range(num_layers)]
    b = latent.shape[0]
    out = sd[prefix + "input.input"].to(dt).repeat(b, 1, 1, 1)
    out = styled_conv(out, latent[:, :, 0], mask, sd, prefix + "conv1.", upsample=False, mask_op=True, noise=noise[0])
    skip = to_rgb(out, latent[:, :, 1], mask, None, sd, prefix + "to_rgb1.", mask_op=True)
    inter = None
    i = 1
    for j, res_log in enumerate(range(3, log_size + 1)):
        conv_masked = not (res_log > 2 + rl // 2)                      # model.py:560,568
        rgb_masked = not (rl != 17 and res_log >= 2 + rl // 2)          # model.py:576
        p1, p2, pr = f"{prefix}convs.{2 * j}.", f"{prefix}convs.{2 * j + 1}.", f"{prefix}to_rgbs.{j}."
        n1, n2 = noise[1 + 2 * j], noise[2 + 2 * j]
        if i < rl:                                                      # model.py:670-684
            out = styled_conv(out, latent[:, :, i], mask, sd, p1, upsample=True, mask_op=conv_masked, noise=n1)
            if i + 2 == split_layer_idx:
                inter = out
            out = styled_conv(out, latent[:, :, i + 1], mask, sd, p2, upsample=False, mask_op=conv_masked, noise=n2)
            st = latent[:, :, i + 2] if (rl == 17 or i + 2 != rl) else latent[:, 0, i + 2]
            skip = to_rgb(out, st, mask, skip, sd, pr, mask_op=rgb_masked)
        else:                                                           # model.py:685-688
            out = styled_conv(out, latent[:, 0, i], mask, sd, p1, upsample=True, mask_op=conv_masked, noise=n1)
            out = styled_conv(out, latent[:, 0, i + 1], mask, sd, p2, upsample=False, mask_op=conv_masked, noise=n2)
            skip = to_rgb(out, latent[:, 0, i + 2], mask, skip, sd, pr, mask_op=rgb_masked)
        i += 2
    return skip, inter


def style_mapping(sd: SD, z: torch.Tensor, n_mlp: int = 8, lr_mlp: float = 0.01, prefix: str = "") -> torch.Tensor:
    """Generator.style: PixelNorm + n_mlp EqualLinear(fused_lrelu) (model.py:14-19,499-509)."""
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(n_mlp):
        x = equal_linear(x, sd[f"{prefix}style.{i + 1}.weight"], sd[f"{prefix}style.{i + 1}.bias"], lr_mlp, True)
    return x


# --------------------------------------------------------------------------------------
# Regional style encoder + Net3 (models/encoders/*, models/networks.py)
# --------------------------------------------------------------------------------------

ENCODER_BLOCKS = [(64, 128, 3), (128, 256, 4), (256, 512, 14), (512, 512, 3)]   # psp_encoders.py:323-328


def encoder_units() -> List[Tuple[int, int, int]]:
    """(in_channel, depth, stride) of the 24 units (helpers.py:25-26)."""
    units = []
    for cin, depth, n in ENCODER_BLOCKS:
        units.append((cin, depth, 2))
        units += [(depth, depth, 1)] * (n - 1)
    return units


def _inorm(x):   # InstanceNorm2d default: affine=False, biased variance, eps=1e-5
    return F.instance_norm(x, eps=1e-5)


def ir_se_unit(x: torch.Tensor, sd: SD, p: str, cin: int, depth: int, stride: int) -> torch.Tensor:
    """models/encoders/helpers.py:122-144 (bottleneck_IR_SE_Ours) + SEModule :56-72."""
    dt = x.dtype
    if cin == depth:
        sc = x[:, :, ::stride, ::stride]                                # MaxPool2d(1, stride)
    else:
        sc = _inorm(F.conv2d(x, sd[p + "shortcut_layer.0.weight"].to(dt), stride=stride))
    r = _inorm(x)
    r = F.conv2d(r, sd[p + "res_layer.1.weight"].to(dt), padding=1)
    r = F.prelu(r, sd[p + "res_layer.2.weight"].to(dt))
    r = F.conv2d(r, sd[p + "res_layer.3.weight"].to(dt), stride=stride, padding=1)
    r = _inorm(r)
    g = r.mean(dim=(2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, sd[p + "res_layer.5.fc1.weight"].to(dt)))
    g = torch.sigmoid(F.conv2d(g, sd[p + "res_layer.5.fc2.weight"].to(dt)))
    return r * g + sc


def masked_region_mean(feats: torch.Tensor, segmap: torch.Tensor) -> torch.Tensor:
    """psp_encoders.py:355-375: per (sample, region) mean of feats over pixels where the
    nearest-resized mask is non-zero; empty region -> zeros.  Vectorised restatement."""
    seg = nearest_resize(segmap.to(feats.dtype), tuple(feats.shape[2:])) != 0   # [B,K,H,W] bool
    segf = seg.to(feats.dtype)
    area = segf.sum(dim=(2, 3))                                           # [B,K]
    tot = torch.einsum("bkhw,bchw->bkc", segf, feats)
    return torch.where(area[..., None] > 0, tot / area.clamp(min=1)[..., None], torch.zeros_like(tot))


def fs_encoder_psp(sd: SD, x: torch.Tensor, segmap: torch.Tensor, prefix: str = "") -> Tuple[torch.Tensor, torch.Tensor]:
    """models/encoders/psp_encoders.py:377-401 (FSEncoder_PSP.forward)."""
    dt = x.dtype
    x = F.conv2d(x, sd[prefix + "input_layer.0.weight"].to(dt), padding=1)
    x = F.prelu(_inorm(x), sd[prefix + "input_layer.2.weight"].to(dt))
    taps = {}
    for i, (cin, depth, stride) in enumerate(encoder_units()):
        x = ir_se_unit(x, sd, f"{prefix}body.{i}.", cin, depth, stride)
        if i in (6, 20, 23):
            taps[i] = x
    codes = torch.cat([masked_region_mean(taps[i], segmap) for i in (6, 20, 23)], dim=2)
    return codes, torch.zeros_like(x)


def local_mlps(sd: SD, vectors: torch.Tensor, prefix: str = "MLPs.") -> torch.Tensor:
    """models/networks.py:23-49 + :223-229: one LocalMLP per region; nn.LeakyReLU() slope 0.01."""
    outs = []
    for i in range(vectors.shape[1]):
        h = equal_linear(vectors[:, i], sd[f"{prefix}{i}.mlp.0.weight"], sd[f"{prefix}{i}.mlp.0.bias"])
        h = F.leaky_relu(h, 0.01)
        h = equal_linear(h, sd[f"{prefix}{i}.mlp.2.weight"], sd[f"{prefix}{i}.mlp.2.bias"])
        outs.append(h.reshape(h.shape[0], -1, 512))
    return torch.stack(outs, dim=1)


def net3_style_codes(sd: SD, vectors: torch.Tensor, latent_avg: Optional[torch.Tensor], remaining_layer_idx: int = 13,
                     start_from_latent_avg: bool = True) -> torch.Tensor:
    """models/networks.py:223-253 (cal_style_codes), learn_in_w=False branch."""
    codes = local_mlps(sd, vectors)
    if not start_from_latent_avg:
        return codes
    b, k = codes.shape[:2]
    la = latent_avg.to(codes.dtype)
    if remaining_layer_idx != 17:
        codes = codes + la[:remaining_layer_idx][None, None]
        rest = la[remaining_layer_idx:][None, None].expand(b, k, -1, -1)
        return torch.cat([codes, rest], dim=2)
    return codes + la[None, None]


def net3_style_vectors(sd: SD, img: torch.Tensor, mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/networks.py:206-221: 1024->256 bilinear (align_corners=False, no antialias) + encoder."""
    return fs_encoder_psp(sd, F.interpolate(img, (256, 256), mode="bilinear"), mask, prefix="encoder.")


def net3_forward(sd: SD, img, mask, latent_avg, *, out_size: int = 1024, remaining_layer_idx: int = 13,
                 noise: Optional[List[torch.Tensor]] = None):
    """models/networks.py:98-159 (Net3.forward, randomize_noise=False)."""
    vec, _ = net3_style_vectors(sd, img, mask)
    codes = net3_style_codes(sd, vec, latent_avg, remaining_layer_idx)
    image, inter = generator_forward(sd, out_size, codes, mask, noise, split_layer_idx=5,
                                     remaining_layer_idx=remaining_layer_idx, prefix="G.")
    return image, inter, codes, vec


# --------------------------------------------------------------------------------------
# BiSeNet face parsing (swap_face_fine/face_parsing/{model,resnet,face_parsing_demo}.py)
# --------------------------------------------------------------------------------------


def _bn(x, sd: SD, p: str):
    """eval-mode BatchNorm2d (running stats, eps 1e-5)."""
    dt = x.dtype
    return F.batch_norm(x, sd[p + "running_mean"].to(dt), sd[p + "running_var"].to(dt),
                        sd[p + "weight"].to(dt), sd[p + "bias"].to(dt), False, 0.0, 1e-5)


def _cbr(x, sd: SD, p: str, stride=1, padding=1):
    """ConvBNReLU, model.py:20-35."""
    return F.relu(_bn(F.conv2d(x, sd[p + "conv.weight"].to(x.dtype), stride=stride, padding=padding), sd, p + "bn."))


def _basic_block(x, sd: SD, p: str, stride: int):
    """resnet.py:21-49."""
    dt = x.dtype
    r = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"].to(dt), stride=stride, padding=1), sd, p + "bn1."))
    r = _bn(F.conv2d(r, sd[p + "conv2.weight"].to(dt), padding=1), sd, p + "bn2.")
    sc = x
    if (p + "downsample.0.weight") in sd:
        sc = _bn(F.conv2d(x, sd[p + "downsample.0.weight"].to(dt), stride=stride), sd, p + "downsample.1.")
    return F.relu(sc + r)


def resnet18_feats(x, sd: SD, p: str = "cp.resnet."):
    """resnet.py:72-81."""
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"].to(x.dtype), stride=2, padding=3), sd, p + "bn1."))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _basic_block(x, sd, f"{p}layer{li}.0.", stride)
        x = _basic_block(x, sd, f"{p}layer{li}.1.", 1)
        feats.append(x)
    return feats[1], feats[2], feats[3]


def _arm(x, sd: SD, p: str):
    """AttentionRefinementModule, model.py:82-89."""
    feat = _cbr(x, sd, p + "conv.")
    a = F.avg_pool2d(feat, feat.shape[2:])
    a = torch.sigmoid(_bn(F.conv2d(a, sd[p + "conv_atten.weight"].to(x.dtype)), sd, p + "bn_atten."))
    return feat * a


def bisenet_forward(sd: SD, x: torch.Tensor):
    """model.py:247-260 (+ContextPath :110-131, FFM :206-216, BiSeNetOutput :50-53).
    Returns the three logits maps at input resolution."""
    h, w = x.shape[2:]
    f8, f16, f32 = resnet18_feats(x, sd)
    avg = _cbr(F.avg_pool2d(f32, f32.shape[2:]), sd, "cp.conv_avg.", padding=0)
    avg_up = nearest_resize(avg, tuple(f32.shape[2:]))
    s32 = _arm(f32, sd, "cp.arm32.") + avg_up
    up32 = _cbr(nearest_resize(s32, tuple(f16.shape[2:])), sd, "cp.conv_head32.")
    s16 = _arm(f16, sd, "cp.arm16.") + up32
    up16 = _cbr(nearest_resize(s16, tuple(f8.shape[2:])), sd, "cp.conv_head16.")
    fcat = torch.cat([f8, up16], dim=1)
    feat = _cbr(fcat, sd, "ffm.convblk.", padding=0)
    a = F.avg_pool2d(feat, feat.shape[2:])
    a = F.relu(F.conv2d(a, sd["ffm.conv1.weight"].to(x.dtype)))
    a = torch.sigmoid(F.conv2d(a, sd["ffm.conv2.weight"].to(x.dtype)))
    fuse = feat * a + feat
    outs = []
    for head, src in (("conv_out.", fuse), ("conv_out16.", up16), ("conv_out32.", up32)):
        y = F.conv2d(_cbr(src, sd, head + "conv."), sd[head + "conv_out.weight"].to(x.dtype))
        outs.append(F.interpolate(y, (h, w), mode="bilinear", align_corners=True))
    return tuple(outs)


def bicubic_taps(factor: int, a: float = -0.5) -> torch.Tensor:
    """face_parsing_demo.py:16-35: 4*factor taps sampled from the cubic kernel, normalised."""
    size = factor * 4
    ks = []
    for i in range(size):
        t = abs((i - math.floor(size / 2) + 0.5) / factor)
        if t <= 1.0:
            v = (a + 2.0) * t ** 3 - (a + 3.0) * t ** 2 + 1
        elif t < 2.0:
            v = a * t ** 3 - 5.0 * a * t ** 2 + 8.0 * a * t - 4.0 * a
        else:
            v = 0.0
        ks.append(v)
    k = torch.tensor(ks, dtype=torch.float32)
    return k / k.sum()


def bicubic_downsample(x: torch.Tensor, factor: int) -> torch.Tensor:
    """face_parsing_demo.py:46-84: reflect pad, vertical then horizontal strided 1-D filter."""
    k = bicubic_taps(factor).to(x)
    n = k.numel()
    c = x.shape[1]
    pad = n - factor
    lo, hi = pad // 2, pad - pad // 2
    x = F.pad(x, (0, 0, lo, hi), mode="reflect")
    x = F.conv2d(x, k.reshape(1, 1, n, 1).repeat(c, 1, 1, 1), stride=(factor, 1), groups=c)
    x = F.pad(x, (lo, hi, 0, 0), mode="reflect")
    return F.conv2d(x, k.reshape(1, 1, 1, n).repeat(c, 1, 1, 1), stride=(1, factor), groups=c)


SEG_MEAN = (0.485, 0.456, 0.406)      # face_parsing/model.py:15
SEG_STD = (0.229, 0.224, 0.225)       # face_parsing/model.py:16


def parser_preprocess(img01: torch.Tensor, size: int = 1024) -> torch.Tensor:
    """face_parsing_demo.py:151-160 for inputs >= 512: bicubic down to 512, clamp, normalise.
    img01 is the ToTensor() image batch [B,3,size,size] in [0,1]."""
    mean = torch.tensor(SEG_MEAN, dtype=img01.dtype, device=img01.device).reshape(1, 3, 1, 1)
    std = torch.tensor(SEG_STD, dtype=img01.dtype, device=img01.device).reshape(1, 3, 1, 1)
    return (bicubic_downsample(img01, size // 512).clamp(0, 1) - mean) / std


# 19-class face-parsing ids -> 12 E4S regions (datasets/dataset.py:58-108)
SEG19_TO_SEG12 = np.zeros(256, dtype=np.uint8)
for _src, _dst in ((12, 1), (13, 1), (2, 2), (3, 2), (4, 3), (5, 3), (17, 4), (10, 5), (1, 6), (7, 7), (8, 7),
                   (14, 8), (11, 9), (6, 10), (9, 11)):
    SEG19_TO_SEG12[_src] = _dst


def face_parse(sd: SD, img01: torch.Tensor, convert_to_seg12: bool = True) -> np.ndarray:
    """face_parsing_demo.py:162-176,187-200 batched: labels uint8 [B,512,512]."""
    logits = bisenet_forward(sd, parser_preprocess(img01, img01.shape[-1]))[0]
    seg = torch.argmax(logits, dim=1).cpu().numpy().astype(np.uint8)
    return SEG19_TO_SEG12[seg] if convert_to_seg12 else seg


def label_to_onehot(label: torch.Tensor, num_cls: int) -> torch.Tensor:
    """utils/torch_utils.py:207-213 (labelMap2OneHot): [B,1,H,W] int64 -> [B,num_cls,H,W] float."""
    b, _, h, w = label.shape
    return torch.zeros(b, num_cls, h, w, device=label.device).scatter_(1, label, 1.0)


def swap_comp_style_vector(sv1: torch.Tensor, sv2: torch.Tensor, comp_indices, below_face_interpolation: bool = False) -> torch.Tensor:
    """swap_face_fine/swap_face_mask.py:336-367 restated per sample (the reference is batch 1: its
    `torch.sum(style_vectors2[:, 9, :]) == 0` test then coincides with the per-sample test used here)."""
    out = sv1.clone()
    for c in comp_indices:                                   # :347-348
        out[:, c, :] = sv2[:, c, :]
    out[:, 7, :] = (sv1[:, 7, :] + sv2[:, 7, :]) / 2         # :354 ears: always the average
    out[:, 11, :] = sv1[:, 11, :]                            # :357 ear-rings: always the target's
    if below_face_interpolation:                             # :360-361
        out[:, 8, :] = (sv1[:, 8, :] + sv2[:, 8, :]) / 2
    empty = sv2[:, 9, :].sum(dim=1) == 0                     # :364-365 source without a mouth region -> target's vector
    out[empty, 9, :] = sv1[empty, 9, :]
    return out


def to_tensor_normalize(img_u8: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> Tuple[torch.Tensor, torch.Tensor]:
    """TO_TENSOR and Compose([TO_TENSOR, NORMALIZE]) (datasets/dataset.py:45, face_swap_video_pipeline.py:338-346), batched:
    uint8 HWC [B,H,W,3] -> (x01 = v/255, (x01 - mean)/std), fp32 NCHW -- torchvision's `.to(float32).div(255)` and `.sub_(mean).div_(std)`."""
    x01 = img_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32, device=img_u8.device).view(1, 3, 1, 1)
    sd = torch.tensor(std, dtype=torch.float32, device=img_u8.device).view(1, 3, 1, 1)
    return x01.contiguous(), x01.sub(m).div(sd).contiguous()


def tensor2im_u8(var: torch.Tensor, is_zero_center: bool = True) -> torch.Tensor:
    """utils/torch_utils.py:64-76 per sample, batched: [B,3,H,W] float -> uint8 [B,H,W,3] (float32 numpy arithmetic)."""
    v = var.detach().cpu().float().permute(0, 2, 3, 1).numpy().copy()
    if is_zero_center:
        v = (v + 1) / 2          # :71
    v[v < 0] = 0                 # :72
    v[v > 1] = 1                 # :73
    v = v * 255                  # :74
    return torch.from_numpy(v.astype("uint8"))


def morphology(x: torch.Tensor, kernel: torch.Tensor, dilate: bool, structuring_element=None, origin=None,
               border_type: str = "geodesic", border_value: float = 0.0, max_val: float = 1e4) -> torch.Tensor:
    """utils/morphology.py:23-108 (dilation) / :111-200 (erosion), the `unfold` engine restated with explicit loops over the
    structuring element (the `convolution` engine computes the same numbers)."""
    se_h, se_w = kernel.shape
    if origin is None:
        origin = [se_h // 2, se_w // 2]                                  # :70-72
    if border_type == "geodesic":                                        # :76-78 / :167-169
        border_value = -max_val if dilate else max_val
    pad = [origin[1], se_w - origin[1] - 1, origin[0], se_h - origin[0] - 1]
    p = F.pad(x.float(), pad, mode="constant", value=border_value)
    nb = torch.zeros_like(kernel, dtype=torch.float32) if structuring_element is None else structuring_element.clone().float()
    nb[kernel == 0] = -max_val                                           # :84-89
    h, w = x.shape[-2:]
    out = None
    for i in range(se_h):
        for j in range(se_w):
            win = p[..., i:i + h, j:j + w]
            if dilate:
                t = win + nb[se_h - 1 - i, se_w - 1 - j]                 # :93-95: max(unfolded + neighborhood.flip)
                out = t if out is None else torch.maximum(out, t)
            else:
                t = win - nb[i, j]                                       # :184-186: min(unfolded - neighborhood)
                out = t if out is None else torch.minimum(out, t)
    return out


# --------------------------------------------------------------------------------------
# paste-back: SoftErosion and the Laplacian-pyramid blend
# --------------------------------------------------------------------------------------

def soft_erosion_kernel(kernel_size: int = 15) -> torch.Tensor:
    """utils/paste_back_tricks.py:24-30 (= gradio_utils/face_swapping.py:31-38): cone kernel max(dist) - dist, normalised."""
    r = kernel_size // 2
    ax = torch.arange(0., kernel_size)
    y, x = torch.meshgrid(ax, ax, indexing="ij")
    dist = torch.sqrt((x - r) ** 2 + (y - r) ** 2)
    k = dist.max() - dist
    return k / k.sum()


def soft_erosion(x: torch.Tensor, kernel_size: int = 15, threshold: float = 0.6, iterations: int = 1):
    """utils/paste_back_tricks.py:32-42 (SoftErosion.forward): depthwise conv (zero padding), optional min-iterations, threshold to 1,
    the rest divided by its GLOBAL maximum.  -> (x, mask bool)."""
    w = soft_erosion_kernel(kernel_size).to(x.device).view(1, 1, kernel_size, kernel_size).repeat(x.shape[1], 1, 1, 1)
    x = x.float()
    pad = kernel_size // 2
    for _ in range(iterations - 1):
        x = torch.min(x, F.conv2d(x, w, groups=x.shape[1], padding=pad))
    x = F.conv2d(x, w, groups=x.shape[1], padding=pad)
    mask = x >= threshold
    x = x.clone()
    x[mask] = 1.0
    x[~mask] /= x[~mask].max()
    return x, mask


def _reflect101(i: int, n: int) -> int:
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def pyr_down(a: np.ndarray, round_u8: bool = False) -> np.ndarray:
    """cv2.pyrDown (multi_band_blending.py:17-19) for float32 HxW[xC] arrays: separable [1 4 6 4 1]/16, BORDER_REFLECT_101, even samples,
    output ((h+1)//2, (w+1)//2).  round_u8: the uint8 form ((sum + 128) >> 8 on integer pixels), what cv2 does when the pipelines hand
    it uint8 images.  Checked against cv2 in oracle/make_golden_r2.py."""
    a = np.asarray(a, np.float32)
    h, w = a.shape[:2]
    oh, ow = (h + 1) // 2, (w + 1) // 2
    row = np.zeros((h, ow) + a.shape[2:], np.float32)
    for x in range(ow):
        xs = [_reflect101(2 * x + i, w) for i in range(-2, 3)]
        row[:, x] = a[:, xs[2]] * 6 + (a[:, xs[1]] + a[:, xs[3]]) * 4 + a[:, xs[0]] + a[:, xs[4]]
    out = np.zeros((oh, ow) + a.shape[2:], np.float32)
    for y in range(oh):
        ys = [_reflect101(2 * y + i, h) for i in range(-2, 3)]
        out[y] = row[ys[2]] * 6 + (row[ys[1]] + row[ys[3]]) * 4 + row[ys[0]] + row[ys[4]]
    if round_u8:
        return np.floor((out + 128.0) / 256.0).astype(np.float32)
    return out * np.float32(1 / 256)


def pyr_up(a: np.ndarray) -> np.ndarray:
    """cv2.pyrUp (multi_band_blending.py:30-31,47) for float32 arrays: zero-insert x2, [1 4 6 4 1]/8 per axis; the left / top neighbour
    of the first sample is reflected (101), the right / bottom neighbour of the last one replicated -- as OpenCV does."""
    def up1(t, axis):
        t = np.moveaxis(t, axis, 0)
        n = t.shape[0]
        out = np.zeros((2 * n,) + t.shape[1:], np.float32)
        for x in range(n):
            left = t[x - 1] if x > 0 else t[min(1, n - 1)]
            right = t[x + 1] if x + 1 < n else t[x]
            out[2 * x] = left + t[x] * 6 + right
            out[2 * x + 1] = (t[x] + right) * 4
        return np.moveaxis(out, 0, axis)
    return up1(up1(np.asarray(a, np.float32), 1), 0) * np.float32(1 / 64)


def laplacian_pyramid_blend(A: np.ndarray, B: np.ndarray, m: np.ndarray, num_levels: int = 6) -> np.ndarray:
    """swap_face_fine/multi_band_blending.py:6-49 (Laplacian_Pyramid_Blending_with_mask) on HxWxC arrays.  uint8 A / B keep cv2's
    per-level integer rounding in their Gaussian pyramids (the pipelines pass np.array(PIL) for A)."""
    def gauss(x):
        u8 = np.asarray(x).dtype == np.uint8
        g = np.asarray(x, np.float32)
        out = [g]
        for _ in range(num_levels):
            g = pyr_down(g, round_u8=u8)
            out.append(g)
        return out
    gpA, gpB, gpM = gauss(A), gauss(B), gauss(m)
    lpA, lpB, gpMr = [gpA[num_levels - 1]], [gpB[num_levels - 1]], [gpM[num_levels - 1]]
    for i in range(num_levels - 1, 0, -1):
        lpA.append(gpA[i - 1] - pyr_up(gpA[i]))
        lpB.append(gpB[i - 1] - pyr_up(gpB[i]))
        gpMr.append(gpM[i - 1])
    LS = [la * gm + lb * (1.0 - gm) for la, lb, gm in zip(lpA, lpB, gpMr)]
    ls = LS[0]
    for i in range(1, num_levels):
        ls = pyr_up(ls) + LS[i]
    return ls


def blending(full_img: np.ndarray, ori_img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """multi_band_blending.py:52-74 for 1024x1024 inputs (the cv2.resize calls are identities there): 10-level blend, clip, uint8."""
    assert full_img.shape[:2] == (1024, 1024) and ori_img.shape[:2] == (1024, 1024)
    img = laplacian_pyramid_blend(full_img, ori_img, np.float32(mask), 10)
    return np.uint8(np.clip(img, 0, 255))

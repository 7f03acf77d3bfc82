"""Deterministic synthetic weights / inputs shaped like the reference checkpoints.

No pretrained E4S / BiSeNet weights are reachable offline, so every parity test, golden
fixture and benchmark uses synthetic parameters.  Each tensor is drawn from a numpy PCG64
stream keyed by (seed, crc32(name)), so the values do not depend on dict order, torch
version or device, and the reference modules, the oracle and the CUDA path all see the
same numbers.  The distributions exercise every epilogue term (the reference's default
init leaves noise weights and biases at zero, which would hide bugs).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Mapping, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def _fan_in(shape) -> int:
    n = 1
    for d in shape[1:]:
        n *= int(d)
    return max(n, 1)


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Rule-based deterministic fill for a parameter/buffer called `name`."""
    g = _rng(seed, name)
    shape = tuple(int(s) for s in shape)
    n = lambda std=1.0: (g.standard_normal(shape) * std).astype(np.float32)
    u = lambda lo, hi: g.uniform(lo, hi, shape).astype(np.float32)
    leaf = name.split(".")[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        a = n(0.1)
    elif leaf == "running_var":
        a = u(0.5, 1.5)
    elif ".noises." in "." + name or name.startswith("noises."):
        a = n()
    elif name.endswith("noise.weight"):
        a = u(0.05, 0.15)
    elif name.endswith("activate.bias") or (leaf == "bias" and "to_rgb" in name and "modulation" not in name):
        a = n(0.1)
    elif name.endswith("modulation.bias"):
        a = 1.0 + n(0.05)
    elif leaf == "kernel":                       # FIR buffers are rebuilt by the module, keep caller's
        raise KeyError(name)
    elif name.endswith("input.input"):
        a = n()
    elif ".bn" in name or "bn_atten" in name or ".downsample.1" in name or leaf in ("bn",):
        a = u(0.5, 1.5) if leaf == "weight" else n(0.1)
    elif leaf == "weight" and len(shape) == 1:   # PReLU slopes
        a = 0.25 + n(0.05)
    elif leaf == "bias":
        a = n(0.1)
    elif leaf == "weight" and (name.startswith("G.") or "conv.weight" in name and len(shape) == 5
                               or ".modulation." in name or name.startswith("style.") or ".mlp." in name
                               or name.startswith("MLPs.")):
        a = n()                                   # StyleGAN-style N(0,1) (equalised lr)
    elif leaf == "weight" and len(shape) == 5:
        a = n()
    elif leaf == "weight":
        a = n((2.0 / _fan_in(shape)) ** 0.5)      # kaiming-normal for ordinary convs
    else:
        a = n()
    return torch.from_numpy(np.ascontiguousarray(a))


def fill_state_dict(shapes: Mapping[str, Iterable[int]], seed: int = 0, skip_suffix=("kernel",)) -> Dict[str, torch.Tensor]:
    """Synthetic state dict for the given {name: shape}; FIR `kernel` buffers are skipped
    (load with strict=False or merge into module.state_dict())."""
    out = {}
    for name, shape in shapes.items():
        if name.split(".")[-1] in skip_suffix:
            continue
        out[name] = synth_tensor(name, tuple(shape), seed)
    return out


def synth_module_weights(module: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Fill `module` in place with synthetic weights; returns the full state dict."""
    sd = module.state_dict()
    new = fill_state_dict({k: v.shape for k, v in sd.items()}, seed)
    for k, v in new.items():
        sd[k] = v.to(sd[k].dtype)
    module.load_state_dict(sd)
    return module.state_dict()


def blocky_labels(batch: int, num_cls: int, size: int, cells: int = 32, seed: int = 0) -> torch.Tensor:
    """FFHQ-like label maps: random class per cell of a cells x cells grid, nearest-upsampled
    (SURVEY.md section 8d config 2).  int64 [B,1,size,size]."""
    g = _rng(seed, f"labels{batch}x{size}x{cells}")
    lab = g.integers(0, num_cls, (batch, 1, cells, cells))
    rep = size // cells
    lab = np.repeat(np.repeat(lab, rep, axis=2), rep, axis=3)
    return torch.from_numpy(lab.astype(np.int64))


def onehot(labels: torch.Tensor, num_cls: int) -> torch.Tensor:
    b, _, h, w = labels.shape
    return torch.zeros(b, num_cls, h, w).scatter_(1, labels, 1.0)


def randn(name: str, shape, seed: int = 0, std: float = 1.0) -> torch.Tensor:
    return torch.from_numpy((_rng(seed, name).standard_normal(tuple(shape)) * std).astype(np.float32))


def smooth_image(name: str, batch: int, size: int, seed: int = 0) -> torch.Tensor:
    """FFHQ-shaped stand-in: low-pass filtered uniform noise in [-1, 1], [B,3,size,size]."""
    g = _rng(seed, name)
    low = g.uniform(-1, 1, (batch, 3, size // 8, size // 8)).astype(np.float32)
    x = torch.nn.functional.interpolate(torch.from_numpy(low), size=(size, size), mode="bilinear", align_corners=False)
    fine = torch.from_numpy(g.uniform(-0.15, 0.15, (batch, 3, size, size)).astype(np.float32))
    return (x + fine).clamp(-1, 1)


def noise_labels(batch: int, num_cls: int, size: int, seed: int = 0) -> torch.Tensor:
    """Adversarial masks: an independent random class per PIXEL (second run): every tile of every
    masked layer sees all regions.  int64 [B,1,size,size]."""
    g = _rng(seed, f"noiselabels{batch}x{size}")
    return torch.from_numpy(g.integers(0, num_cls, (batch, 1, size, size)).astype(np.int64))


# (region id, centre x, centre y, radius x, radius y, tilt in degrees) in face coordinates [-1, 1]^2, painted in this order.
# Region ids follow datasets/dataset.py:58-108 (0 background, 1 lips, 2 brows, 3 eyes, 4 hair, 5 nose, 6 skin, 7 ears, 8 neck,
# 9 mouth, 10 glasses, 11 ear-rings).
_FACE_PARTS = (
    (4, 0.00, -0.20, 0.70, 0.80, 0.0),     # hair
    (8, 0.00, 0.80, 0.30, 0.42, 0.0),      # neck
    (7, -0.50, 0.02, 0.08, 0.16, 8.0), (7, 0.50, 0.02, 0.08, 0.16, -8.0),      # ears
    (11, -0.52, 0.24, 0.035, 0.06, 0.0), (11, 0.52, 0.24, 0.035, 0.06, 0.0),   # ear-rings
    (6, 0.00, 0.06, 0.46, 0.60, 0.0),      # skin
    (2, -0.21, -0.21, 0.13, 0.030, 8.0), (2, 0.21, -0.21, 0.13, 0.030, -8.0),  # brows
    (10, -0.21, -0.09, 0.14, 0.085, 0.0), (10, 0.21, -0.09, 0.14, 0.085, 0.0), # glasses (half of the samples)
    (3, -0.21, -0.09, 0.085, 0.040, 0.0), (3, 0.21, -0.09, 0.085, 0.040, 0.0), # eyes
    (5, 0.00, 0.10, 0.075, 0.16, 0.0),     # nose
    (1, 0.00, 0.38, 0.16, 0.065, 0.0),     # lips
    (9, 0.00, 0.38, 0.095, 0.022, 0.0),    # mouth interior
)


def face_labels(batch: int, size: int = 512, seed: int = 0) -> torch.Tensor:
    """Procedural face-shaped label maps with CURVED region boundaries (ellipses with per-sample shift, scale and tilt; the bundled
    CelebA-HQ masks of the reference are not on the GPU box): 12 regions with realistic area fractions (hair / skin / background
    dominate, brows / eyes / lips are small), so masked layers see 1-3 regions per 16x8 tile along curved boundaries instead of the
    tile-aligned cells of blocky_labels.  int64 [B,1,size,size]."""
    g = _rng(seed, f"facelabels{batch}x{size}")
    ys, xs = np.meshgrid(np.linspace(-1, 1, size, dtype=np.float32), np.linspace(-1, 1, size, dtype=np.float32), indexing="ij")
    out = np.zeros((batch, 1, size, size), dtype=np.int64)
    for b in range(batch):
        dx, dy = g.uniform(-0.08, 0.08, 2)
        sc = g.uniform(1.15, 1.45)
        rot = np.deg2rad(g.uniform(-10, 10))
        glasses, rings = g.random() < 0.5, g.random() < 0.5
        c, s = np.cos(rot), np.sin(rot)
        fx = ((xs - dx) * c + (ys - dy) * s) / sc           # image -> face coordinates
        fy = (-(xs - dx) * s + (ys - dy) * c) / sc
        lab = out[b, 0]
        for rid, cx, cy, rx, ry, tilt in _FACE_PARTS:
            if (rid == 10 and not glasses) or (rid == 11 and not rings):
                continue
            t = np.deg2rad(tilt)
            ux = (fx - cx) * np.cos(t) + (fy - cy) * np.sin(t)
            uy = -(fx - cx) * np.sin(t) + (fy - cy) * np.cos(t)
            lab[(ux / rx) ** 2 + (uy / ry) ** 2 <= 1.0] = rid
    return torch.from_numpy(out)


def make_labels(kind: str, batch: int, num_cls: int = 12, size: int = 512, seed: int = 0) -> torch.Tensor:
    """The three mask families of the benchmark / parity tests: "blocky" (32x32 random cells), "noise" (per pixel), "face" (curved)."""
    if kind == "blocky":
        return blocky_labels(batch, num_cls, size, cells=32, seed=seed)
    if kind == "noise":
        return noise_labels(batch, num_cls, size, seed=seed)
    if kind == "face":
        if num_cls < 12:
            raise ValueError("face_labels paints 12 regions")
        return face_labels(batch, size, seed=seed)
    raise ValueError(kind)


def smooth_image_u8(name: str, batch: int, size: int, seed: int = 0) -> torch.Tensor:
    """The same stand-in as the u8 HWC image a pipeline holds before TO_TENSOR: [B,size,size,3] uint8."""
    x = smooth_image(name, batch, size, seed)
    return ((x + 1) / 2).clamp(0, 1).mul(255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()

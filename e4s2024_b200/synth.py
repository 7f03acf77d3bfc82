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

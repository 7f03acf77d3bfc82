"""Style recombination between the regional encoder and the generator -- drop-in for
`swap_face_fine/swap_face_mask.py:336-367` (the step between a9 and a4).

Same signature and rules as the reference; batched (the reference is batch 1) and without the host read-back of
`torch.sum(style_vectors2[:, 9, :]) == 0`: the empty-mouth test runs per sample on the device."""
from __future__ import annotations

import torch

from . import _lib as L


def swap_comp_style_vector(style_vectors1: torch.Tensor, style_vectors2: torch.Tensor, comp_indices=[], belowFace_interpolation=False):
    """style_vectors1 [B,#comp,D] (target), style_vectors2 [B,#comp,D] (source) -> recombined [B,#comp,D]."""
    assert comp_indices is not None
    if style_vectors1.shape != style_vectors2.shape or style_vectors1.dim() != 3:
        raise L.E4SError(f"style vectors must both be [B, #comp, D]: {tuple(style_vectors1.shape)} vs {tuple(style_vectors2.shape)}")
    k = style_vectors1.shape[1]
    mask = 0
    for c in comp_indices:
        c = int(c)
        if not 0 <= c < k:
            raise L.E4SError(f"component index {c} out of range for {k} components")
        mask |= 1 << c
    return L.swap_comp_styles(style_vectors1.contiguous().float(), style_vectors2.contiguous().float(), mask, bool(belowFace_interpolation))

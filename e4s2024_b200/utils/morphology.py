"""Grey-scale morphology of the paste-back masks -- drop-in for `utils/morphology.py:23-200` (first piece).

Same signatures as the reference (`dilation`, `erosion`, `opening`, `closing`); both of its `engine` values compute the same
numbers, so the argument is accepted and ignored.  Borders: `geodesic` (default) and `constant`."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _lib as L

__all__ = ["dilation", "erosion", "opening", "closing"]


def _prepare(tensor, kernel, structuring_element, origin, max_val):
    if not isinstance(tensor, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(tensor)}")
    if len(tensor.shape) != 4:
        raise ValueError(f"Input size must have 4 dimensions. Got {tensor.dim()}")
    if not isinstance(kernel, torch.Tensor):
        raise TypeError(f"Kernel type is not a torch.Tensor. Got {type(kernel)}")
    if len(kernel.shape) != 2:
        raise ValueError(f"Kernel size must have 2 dimensions. Got {kernel.dim()}")
    se_h, se_w = kernel.shape
    if origin is None:
        origin = [se_h // 2, se_w // 2]
    # morphology.py:84-89 / :175-180: 0 (or the non-flat element) inside the kernel's support, -max_val outside
    nb = torch.zeros_like(kernel, dtype=torch.float32) if structuring_element is None else structuring_element.clone().float()
    nb[kernel == 0] = -max_val
    return tensor.contiguous().float(), nb.to(tensor.device).contiguous(), origin


def _border(border_type, border_value, geodesic_value):
    if border_type == "geodesic":
        return geodesic_value
    if border_type == "constant":
        return float(border_value)
    raise NotImplementedError(f"border_type {border_type!r}: the CUDA kernel implements 'geodesic' and 'constant'")


def dilation(tensor: torch.Tensor, kernel: torch.Tensor, structuring_element: Optional[torch.Tensor] = None,
             origin: Optional[List[int]] = None, border_type: str = "geodesic", border_value: float = 0.0, max_val: float = 1e4,
             engine: str = "unfold") -> torch.Tensor:
    """morphology.py:23-108."""
    x, nb, origin = _prepare(tensor, kernel, structuring_element, origin, max_val)
    return L.morphology(x, nb, origin, _border(border_type, border_value, -max_val), True).view_as(tensor)


def erosion(tensor: torch.Tensor, kernel: torch.Tensor, structuring_element: Optional[torch.Tensor] = None,
            origin: Optional[List[int]] = None, border_type: str = "geodesic", border_value: float = 0.0, max_val: float = 1e4,
            engine: str = "unfold") -> torch.Tensor:
    """morphology.py:111-200."""
    x, nb, origin = _prepare(tensor, kernel, structuring_element, origin, max_val)
    return L.morphology(x, nb, origin, _border(border_type, border_value, max_val), False)


def opening(tensor, kernel, structuring_element=None, origin=None, border_type="geodesic", border_value=0.0, max_val=1e4,
            engine="unfold"):
    """dilation(erosion(x)) (morphology.py: opening)."""
    kw = dict(structuring_element=structuring_element, origin=origin, border_type=border_type, border_value=border_value, max_val=max_val)
    return dilation(erosion(tensor, kernel, **kw), kernel, **kw)


def closing(tensor, kernel, structuring_element=None, origin=None, border_type="geodesic", border_value=0.0, max_val=1e4,
            engine="unfold"):
    """erosion(dilation(x)) (morphology.py: closing)."""
    kw = dict(structuring_element=structuring_element, origin=origin, border_type=border_type, border_value=border_value, max_val=max_val)
    return erosion(dilation(tensor, kernel, **kw), kernel, **kw)

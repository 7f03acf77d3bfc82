"""Tensor helpers on either side of the hot path (reference: utils/torch_utils.py:64-76,207-223)."""
import numpy as np
import torch
from PIL import Image

from .. import _lib as L


def labelMap2OneHot(label, num_cls):
    """[bs,1,H,W] integer label map -> one-hot float [bs,num_cls,H,W] (torch_utils.py:207-213)."""
    if label.numel() and (int(label.min()) < 0 or int(label.max()) >= num_cls):       # scatter_ raises here in the reference
        raise L.E4SError(f"labelMap2OneHot: labels must lie in [0, {num_cls})")
    return L.labels_to_onehot(label[:, 0].to(torch.uint8).contiguous(), num_cls)             # CUDA only: no CPU fallback


def tensor2im(var, is_zero_center: bool = True):
    """torch_utils.py:64-76."""
    var = var.squeeze()
    if var.ndim == 3:
        var = var.permute(1, 2, 0)
    var = var.cpu().detach().numpy()
    if is_zero_center:
        var = (var + 1) / 2
    var = np.clip(var, 0, 1) * 255
    return Image.fromarray(var.astype("uint8"))


def tensor2im_batch(var: torch.Tensor, is_zero_center: bool = True) -> torch.Tensor:
    """Batched, on-device form of tensor2im's arithmetic: [B,3,H,W] float (CUDA) -> uint8 [B,H,W,3] (CUDA), byte for byte what
    `tensor2im` puts into the PIL image for each sample; copy it to the host (a quarter of the float bytes) and wrap rows with
    `Image.fromarray` as needed."""
    return L.tensor2im_u8(var.contiguous().float(), is_zero_center)


def remove_module_prefix(state_dict, prefix):
    """torch_utils.py:216-223."""
    return {k.replace(prefix, "", 1): v for k, v in state_dict.items()}

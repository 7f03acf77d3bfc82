"""SoftErosion -- drop-in for `utils/paste_back_tricks.py:17-42` (the same class lives in gradio_utils/face_swapping.py:24-50): softens
the border of the paste-back mask.  Same constructor, `weight` buffer and forward() -> (x, mask); the depthwise cone-kernel convolution,
the min-iterations and the threshold / global-max normalisation run as CUDA kernels (e4s_depthwise_conv_f32, e4s_soft_erosion_finish_f32)."""
import torch
from torch import nn

from .. import _lib as L


class SoftErosion(nn.Module):
    def __init__(self, kernel_size=15, threshold=0.6, iterations=1):
        super().__init__()
        r = kernel_size // 2
        self.padding = r
        self.iterations = iterations
        self.threshold = threshold
        y_indices, x_indices = torch.meshgrid(torch.arange(0., kernel_size), torch.arange(0., kernel_size), indexing="ij")
        dist = torch.sqrt((x_indices - r) ** 2 + (y_indices - r) ** 2)
        kernel = dist.max() - dist
        kernel /= kernel.sum()
        self.register_buffer("weight", kernel.view(1, 1, *kernel.shape))

    @torch.no_grad()
    def forward(self, x):
        """x [B,C,H,W] on the GPU -> (softened mask fp32, hard mask bool); the normalising maximum is taken over the whole tensor, like the
        reference's `x[~mask].max()`."""
        x = x.float().contiguous()
        w = self.weight[0, 0].contiguous().float()
        for _ in range(self.iterations - 1):
            x = L.depthwise_conv(x, w, min_with_input=True)
        x = L.depthwise_conv(x, w)
        return L.soft_erosion_finish(x, float(self.threshold))

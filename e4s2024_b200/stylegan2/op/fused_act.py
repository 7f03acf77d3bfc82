"""fused bias + leaky ReLU (reference: models/stylegan2/op/fused_act.py:72-85, kernel
fused_bias_act_kernel.cu:18-49, act=3 grad=0).  Forward only (inference hot path)."""
import torch
from torch import nn

from ... import _lib as L


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    if not input.is_cuda:                     # the reference op raises on CPU tensors too (fused_bias_act.cpp:13)
        raise RuntimeError("fused_leaky_relu: input must be a CUDA tensor")
    x = input.contiguous().float()
    b = None if bias is None else bias.detach().contiguous().float()
    return L.bias_act(x, b, float(negative_slope), float(scale))


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

"""fused bias + leaky ReLU (reference: models/stylegan2/op/fused_act.py:18-85, kernel fused_bias_act_kernel.cu:18-49).
Differentiable like the reference op, first and second order (the gradient kernel is `e4s_bias_act_grad_f32`)."""
import torch
from torch import nn
from torch.autograd import Function

from ... import _lib as L


class FusedLeakyReLUFunctionBackward(Function):
    """fused_act.py:18-47."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = L.bias_act_grad(grad_output.contiguous().float(), None, out, negative_slope, scale)
        dim = [0]
        if grad_input.ndim > 2:
            dim += list(range(2, grad_input.ndim))
        grad_bias = grad_input.sum(dim).detach()                          # :31-36 (the reference sums with torch too)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gb = None if gradgrad_bias is None else gradgrad_bias.contiguous().float()
        gradgrad_out = L.bias_act_grad(gradgrad_input.contiguous().float(), gb, out, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    """fused_act.py:50-69."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = L.bias_act(input, bias, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    if not input.is_cuda and not getattr(L, "EMULATED", False):   # the reference op raises on CPU tensors too (fused_bias_act.cpp:13)
        raise RuntimeError("fused_leaky_relu: input must be a CUDA tensor")
    x = input.contiguous().float()
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or (bias is not None and bias.requires_grad))
    if not needs_grad:
        b = None if bias is None else bias.detach().contiguous().float()
        return L.bias_act(x, b, float(negative_slope), float(scale))
    if bias is None:
        raise RuntimeError("fused_leaky_relu: the differentiable form needs a bias (as the reference op does)")
    return FusedLeakyReLUFunction.apply(x, bias.contiguous().float(), float(negative_slope), float(scale))


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

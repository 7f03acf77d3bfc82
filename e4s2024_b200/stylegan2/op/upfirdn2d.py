"""upfirdn2d (reference: models/stylegan2/op/upfirdn2d.py:17-147 -> upfirdn2d_kernel.cu:140-272).
Same call signature; NCHW fp32 CUDA tensors.  Differentiable like the reference op (first and second order): the gradient of
an upfirdn2d is the upfirdn2d of the incoming gradient with the flipped kernel, up and down swapped and the pads of
`UpFirDn2d.forward` (:100-105) -- the same CUDA kernel does all three."""
import torch
from torch.autograd import Function

from ... import _lib as L


class UpFirDn2dBackward(Function):
    """upfirdn2d.py:17-85."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        g = grad_output.contiguous().float().reshape(in_size[0], in_size[1], out_size[0], out_size[1])
        grad_input = L.upfirdn2d_general(g, grad_kernel, down[0], down[1], up[0], up[1], *g_pad)
        if tuple(grad_input.shape) != tuple(in_size):
            raise L.E4SError(f"upfirdn2d backward: gradient shape {tuple(grad_input.shape)} != input shape {tuple(in_size)}")
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad, ctx.in_size, ctx.out_size = up, down, pad, in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gg = gradgrad_input.contiguous().float()
        out = L.upfirdn2d_general(gg, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad)
        return out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """upfirdn2d.py:87-139."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = input.shape
        out = L.upfirdn2d_general(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        out_h, out_w = out.shape[2], out.shape[3]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]).contiguous())
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        g_pad_x0 = kernel_w - pad_x0 - 1                                   # :100-105
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad, ctx.in_size,
                                             ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not input.is_cuda and not getattr(L, "EMULATED", False):
        raise RuntimeError("upfirdn2d: input must be a CUDA tensor")
    x = input.contiguous().float()
    k = kernel.detach().to(x.device).contiguous().float()
    if not (torch.is_grad_enabled() and x.requires_grad):
        return L.upfirdn2d(x, k, int(up), int(down), int(pad[0]), int(pad[1]))
    return UpFirDn2d.apply(x, k, (int(up), int(up)), (int(down), int(down)), (int(pad[0]), int(pad[1]), int(pad[0]), int(pad[1])))

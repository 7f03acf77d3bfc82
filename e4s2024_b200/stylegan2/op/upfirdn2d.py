"""upfirdn2d (reference: models/stylegan2/op/upfirdn2d.py:142-147 -> upfirdn2d_kernel.cu:140-272).
Same call signature; NCHW fp32 CUDA tensors; forward only."""
from ... import _lib as L


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not input.is_cuda:
        raise RuntimeError("upfirdn2d: input must be a CUDA tensor")
    x = input.contiguous().float()
    k = kernel.detach().to(x.device).contiguous().float()
    return L.upfirdn2d(x, k, int(up), int(down), int(pad[0]), int(pad[1]))

"""Drop-in for the reference package `models.stylegan2.op` (models/stylegan2/op/__init__.py:1-2):
same names, argument meaning and error behaviour, backed by libe4s_b200.so instead of the
JIT-built pybind extensions.  Like the reference ops these run on CUDA tensors only."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
from . import conv2d_gradfix

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix"]

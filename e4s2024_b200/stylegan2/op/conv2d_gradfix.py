"""conv2d_gradfix: on every torch other than 1.7/1.8 the reference module is a pass-through to
torch.nn.functional (models/stylegan2/op/conv2d_gradfix.py:22-92).  It only matters for training
(double backward); kept so `from models.stylegan2.op import conv2d_gradfix` keeps resolving."""
import contextlib

from torch.nn import functional as F

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return F.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                              output_padding=output_padding, dilation=dilation, groups=groups)

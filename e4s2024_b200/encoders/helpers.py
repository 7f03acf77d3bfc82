"""Building blocks of the regional style encoder -- drop-in for the pieces of
`models/encoders/helpers.py` the hot path instantiates (get_block :25-26, SEModule :56-72,
bottleneck_IR_SE_Ours :122-144).  Parameters keep the reference's names so checkpoints load.

Engine mapping of one bottleneck_IR_SE_Ours unit (NHWC):
    stats(x) -> conv3x3(IN applied while gathering; PReLU in the epilogue) -> conv3x3(stride)
    -> stats -> [1x1 stride-s shortcut conv -> stats] -> one fused  IN(res)*gate + IN?(shortcut)  pass.
The SE gate that follows an affine-free InstanceNorm is the constant 0.5: the global average of an
instance-normalised map is exactly 0, fc1/fc2 have no bias, relu(0)=0 and sigmoid(0)=0.5 (the
reference evaluates the same thing up to ~1e-8 of rounding noise).  SEModule.forward on arbitrary
input still computes the general gate.
"""
from collections import namedtuple

import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..engine import View


class Bottleneck(namedtuple("Block", ["in_channel", "depth", "stride"])):
    """A named tuple describing a ResNet block."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def _ver(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


class PackedConv2d(nn.Conv2d):
    """nn.Conv2d parameter container (bias-free) that also caches its engine packing."""

    def packed(self, cin_pad=None):
        key = _ver(self.weight)
        if getattr(self, "_pack", None) is None or self._pack[0] != key:
            self._pack = (key, E.pack_conv_weight(self.weight.detach().float(), cin_pad=cin_pad))
        return self._pack[1]


class SEModule(nn.Module):
    def __init__(self, channels, reduction):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)

    def gate(self, pooled: torch.Tensor) -> torch.Tensor:
        """pooled [B,C] -> sigmoid(fc2(relu(fc1(pooled))))."""
        c = pooled.shape[1]
        h = L.vec_fc(pooled, self.fc1.weight.detach().reshape(-1, c).contiguous(), None, None, L.ACT_RELU)
        return L.vec_fc(h, self.fc2.weight.detach().reshape(c, -1).contiguous(), None, None, L.ACT_SIGMOID)

    def forward(self, x):
        xn = L.nchw_to_nhwc(x.contiguous().float())
        c = x.shape[1]
        mean, _ = L.chan_stats(xn, c, want_rstd=False)
        return L.nhwc_to_nchw(L.residual_combine(xn, c, gate=self.gate(mean)))


_half_cache = {}


def _half_gate(b, c, device):
    key = (b, c, device)
    if key not in _half_cache:
        _half_cache[key] = torch.full((b, c), 0.5, device=device, dtype=torch.float32)
    return _half_cache[key]


class bottleneck_IR_SE_Ours(nn.Module):
    def __init__(self, in_channel, depth, stride):
        super().__init__()
        self.in_channel, self.depth, self.stride = in_channel, depth, stride
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(PackedConv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.InstanceNorm2d(depth))
        self.res_layer = nn.Sequential(nn.InstanceNorm2d(in_channel),
                                       PackedConv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
                                       nn.PReLU(depth),
                                       PackedConv2d(depth, depth, (3, 3), stride, 1, bias=False),
                                       nn.InstanceNorm2d(depth),
                                       SEModule(depth, 16))

    def run(self, x: View) -> View:
        b, h, w = x.bhw
        stats_x = L.chan_stats(x.t, self.in_channel)
        c1 = E.conv(x, self.res_layer[1].packed(), in_stats=stats_x, act=L.ACT_PRELU,
                    prelu=self.res_layer[2].weight.detach())
        c2 = E.conv(c1, self.res_layer[3].packed(), stride=self.stride)
        stats_c2 = L.chan_stats(c2.t, self.depth)
        gate = _half_gate(b, self.depth, x.t.device)
        if self.in_channel == self.depth:
            out = L.residual_combine(c2.t, self.depth, a_stats=stats_c2, gate=gate, r=x.t, r_sub=self.stride)
        else:
            sc = E.conv(x, self.shortcut_layer[0].packed(), stride=self.stride, pad=0)
            stats_sc = L.chan_stats(sc.t, self.depth)
            out = L.residual_combine(c2.t, self.depth, a_stats=stats_c2, gate=gate, r=sc.t, r_sub=1, r_stats=stats_sc)
        return View(out)

    def forward(self, x):
        return L.nhwc_to_nchw(self.run(View(L.nchw_to_nhwc(x.contiguous().float()))).t)

"""ResNet-18 context backbone of BiSeNet -- drop-in for `swap_face_fine/face_parsing/resnet.py`
(BasicBlock :21-49, Resnet18 :59-81) with the reference's parameter names.  BatchNorm is applied in
eval mode as a per-channel scale/shift in the conv epilogue; the residual add + ReLU are fused too.
Unlike the reference, construction does not download ImageNet weights (resnet.py:83-90): FaceParser
loads the face-parsing checkpoint over them immediately anyway."""
import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..engine import View

RGB_PAD = 8


def _ver(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


def bn_affine(bn: nn.BatchNorm2d):
    """eval-mode BN as y = x*scale + shift (cached per parameter version)."""
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = _ver(*ts)
    c = getattr(bn, "_e4s_affine", None)
    if c is None or c[0] != key:
        scale = (bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)).float().contiguous()
        shift = (bn.bias.detach() - bn.running_mean.detach() * scale).float().contiguous()
        c = (key, scale, shift)
        bn._e4s_affine = c
    return c[1], c[2]


import os
# BiSeNet's label map has to match the reference's argmax, so its convolutions need fp32-class logits:
#   "tc16" (default): tcgen05 with the fp16 hi/lo 3-pass split (operands exact to 2^-22, fp32 accumulate) -- the same
#           kernels as the rest of the path, E4SConv.tc_fmt = E4S_TC_F16;
#   "f32":  the exact-fp32 CUDA-core engine (4-5x slower; the cross-check of the tensor-core mode);
#   "tc":   bf16 hi/lo split (2^-17 operands): logits within 8e-5 relative, 15 of 1,048,576 labels differ -- not the bar.
# E4S_BISENET_ENGINE / set_bisenet_engine select the mode.
_BISENET_ENGINE = [os.environ.get("E4S_BISENET_ENGINE", "tc16")]
_TC_FMT = {"tc": L.TC_BF16, "tc16": L.TC_F16}


def set_bisenet_engine(name: str):
    if name not in ("f32", "tc", "tc16"):
        raise ValueError(name)
    _BISENET_ENGINE[0] = name


def bisenet_engine() -> str:
    return _BISENET_ENGINE[0]


def _conv_engine() -> str:
    """engine.conv's engine name for the current mode."""
    return "f32" if _BISENET_ENGINE[0] == "f32" else "tc"


def packed(conv: nn.Conv2d, cin_pad=None, want_tc=False):
    """Engine packing of a BiSeNet conv, cached per weight version and mode (tensor-core image in the tc modes)."""
    mode = _BISENET_ENGINE[0]
    want_tc = want_tc or mode != "f32"
    key = _ver(conv.weight) + (want_tc, mode)
    c = getattr(conv, "_e4s_pack", None)
    if c is None or c[0] != key:
        c = (key, E.pack_conv_weight(conv.weight.detach().float(), cin_pad=cin_pad, want_tc=want_tc,
                                     tc_fmt=_TC_FMT.get(mode, L.TC_BF16)))
        conv._e4s_pack = c
    return c[1]


def conv_bn(x: View, conv: nn.Conv2d, bn: nn.BatchNorm2d, relu: bool, res: View = None, in_shift=0, out: View = None,
            cin_pad=None) -> View:
    scale, shift = bn_affine(bn)
    return E.conv(x, packed(conv, cin_pad), stride=conv.stride[0], pad=conv.padding[0], in_shift=in_shift, ch_scale=scale,
                  ch_shift=shift, res=res, act=L.ACT_RELU if relu else L.ACT_NONE, out=out, engine=_conv_engine())


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    def __init__(self, in_chan, out_chan, stride=1):
        super().__init__()
        self.conv1 = conv3x3(in_chan, out_chan, stride)
        self.bn1 = nn.BatchNorm2d(out_chan)
        self.conv2 = conv3x3(out_chan, out_chan)
        self.bn2 = nn.BatchNorm2d(out_chan)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if in_chan != out_chan or stride != 1:
            self.downsample = nn.Sequential(nn.Conv2d(in_chan, out_chan, kernel_size=1, stride=stride, bias=False),
                                            nn.BatchNorm2d(out_chan))

    def run(self, x: View, out: View = None) -> View:
        r = conv_bn(x, self.conv1, self.bn1, relu=True)
        sc = x if self.downsample is None else conv_bn(x, self.downsample[0], self.downsample[1], relu=False)
        return conv_bn(r, self.conv2, self.bn2, relu=True, res=sc, out=out)      # relu(shortcut + bn2(conv2(r)))


def create_layer_basic(in_chan, out_chan, bnum, stride=1):
    layers = [BasicBlock(in_chan, out_chan, stride=stride)]
    for _ in range(bnum - 1):
        layers.append(BasicBlock(out_chan, out_chan, stride=1))
    return nn.Sequential(*layers)


class Resnet18(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = create_layer_basic(64, 64, bnum=2, stride=1)
        self.layer2 = create_layer_basic(64, 128, bnum=2, stride=2)
        self.layer3 = create_layer_basic(128, 256, bnum=2, stride=2)
        self.layer4 = create_layer_basic(256, 512, bnum=2, stride=2)

    def run(self, x: View, feat8_out: View = None):
        """x: NHWC with RGB_PAD channels.  feat8 may be written straight into a slice of the FFM concat buffer."""
        y = conv_bn(x, self.conv1, self.bn1, relu=True, cin_pad=RGB_PAD)
        y = View(L.maxpool3x3s2(y.t))
        y = self.layer1[1].run(self.layer1[0].run(y))
        feat8 = self.layer2[1].run(self.layer2[0].run(y), out=feat8_out)
        feat16 = self.layer3[1].run(self.layer3[0].run(feat8))
        feat32 = self.layer4[1].run(self.layer4[0].run(feat16))
        return feat8, feat16, feat32

    def forward(self, x):
        f8, f16, f32 = self.run(View(L.nchw_to_nhwc(x.contiguous().float(), RGB_PAD)))
        return tuple(L.nhwc_to_nchw(f.t, f.c) for f in (f8, f16, f32))

    def get_params(self):
        wd_params, nowd_params = [], []
        for _, module in self.named_modules():
            if isinstance(module, (nn.Linear, nn.Conv2d)):
                wd_params.append(module.weight)
                if module.bias is not None:
                    nowd_params.append(module.bias)
            elif isinstance(module, nn.BatchNorm2d):
                nowd_params += list(module.parameters())
        return wd_params, nowd_params

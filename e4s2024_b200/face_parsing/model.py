"""BiSeNet face parser -- drop-in for `swap_face_fine/face_parsing/model.py` (ConvBNReLU :20-41,
BiSeNetOutput :43-70, AttentionRefinementModule :73-95, ContextPath :98-148, FeatureFusionModule
:186-233, BiSeNet :236-278) with the reference's parameter names (checkpoints load unchanged).
forward(x) -> (out, out16, out32), each [B,n_classes,H,W].  Inference only (BatchNorm in eval form)."""
import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..engine import View
from .resnet import RGB_PAD, Resnet18, bn_affine, conv_bn, packed

seg_mean = torch.tensor([0.485, 0.456, 0.406]).float().reshape(1, 3, 1, 1)       # model.py:15 (CPU copies; moved on use)
seg_std = torch.tensor([0.229, 0.224, 0.225]).float().reshape(1, 3, 1, 1)        # model.py:16


class ConvBNReLU(nn.Module):
    def __init__(self, in_chan, out_chan, ks=3, stride=1, padding=1, *args, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_chan, out_chan, kernel_size=ks, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(out_chan)
        nn.init.kaiming_normal_(self.conv.weight, a=1)

    def run(self, x: View, in_shift=0, out: View = None) -> View:
        return conv_bn(x, self.conv, self.bn, relu=True, in_shift=in_shift, out=out)

    def vec(self, pooled: torch.Tensor) -> torch.Tensor:
        """1x1 ConvBNReLU applied to a pooled [B,C] vector."""
        scale, shift = bn_affine(self.bn)
        w = self.conv.weight.detach().reshape(self.conv.out_channels, -1).contiguous()
        return L.vec_fc(pooled, w, scale, shift, L.ACT_RELU)

    def forward(self, x):
        return L.nhwc_to_nchw(self.run(View(L.nchw_to_nhwc(x.contiguous().float()))).t)


class BiSeNetOutput(nn.Module):
    def __init__(self, in_chan, mid_chan, n_classes, *args, **kwargs):
        super().__init__()
        self.conv = ConvBNReLU(in_chan, mid_chan, ks=3, stride=1, padding=1)
        self.conv_out = nn.Conv2d(mid_chan, n_classes, kernel_size=1, bias=False)
        nn.init.kaiming_normal_(self.conv_out.weight, a=1)

    def run(self, x: View) -> View:
        """-> logits NHWC [B,h,w,pad4(n_classes)] (only the first n_classes channels are meaningful)."""
        mid = self.conv.run(x)
        pw = packed(self.conv_out)
        b, h, w = mid.bhw
        out = View(E.new_nhwc(b, h, w, pw.cout_pad, x.t.device), c=pw.cout)
        return E.conv(mid, pw, pad=0, out=out, engine="f32")


class AttentionRefinementModule(nn.Module):
    def __init__(self, in_chan, out_chan, *args, **kwargs):
        super().__init__()
        self.conv = ConvBNReLU(in_chan, out_chan, ks=3, stride=1, padding=1)
        self.conv_atten = nn.Conv2d(out_chan, out_chan, kernel_size=1, bias=False)
        self.bn_atten = nn.BatchNorm2d(out_chan)
        self.sigmoid_atten = nn.Sigmoid()
        nn.init.kaiming_normal_(self.conv_atten.weight, a=1)

    def run(self, x: View, add: View) -> View:
        """feat * sigmoid(bn(conv1x1(avgpool(feat)))) + nearest_resize(add)   (model.py:82-89 + :122/:127)."""
        feat = self.conv.run(x)
        c = feat.c
        mean, _ = L.chan_stats(feat.t, c, want_rstd=False)
        scale, shift = bn_affine(self.bn_atten)
        atten = L.vec_fc(mean, self.conv_atten.weight.detach().reshape(c, c).contiguous(), scale, shift, L.ACT_SIGMOID)
        return View(L.residual_combine(feat.t, c, gate=atten, r=add.t, r_sub=1))


class ContextPath(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.resnet = Resnet18()
        self.arm16 = AttentionRefinementModule(256, 128)
        self.arm32 = AttentionRefinementModule(512, 128)
        self.conv_head32 = ConvBNReLU(128, 128, ks=3, stride=1, padding=1)
        self.conv_head16 = ConvBNReLU(128, 128, ks=3, stride=1, padding=1)
        self.conv_avg = ConvBNReLU(512, 128, ks=1, stride=1, padding=0)

    def run(self, x: View):
        b, h0, w0 = x.bhw
        if h0 % 32 or w0 % 32:
            raise L.E4SError("BiSeNet input height/width must be multiples of 32")
        dev = x.t.device
        fcat = E.new_nhwc(b, h0 // 8, w0 // 8, 256, dev)            # FFM concat buffer: [feat8 | feat_cp8]
        feat8, feat16, feat32 = self.resnet.run(x, feat8_out=View(fcat, 128, 0))
        avg, _ = L.chan_stats(feat32.t, 512, want_rstd=False)       # F.avg_pool2d(feat32, full)
        avg = self.conv_avg.vec(avg)                                # [B,128]; nearest-upsampled = broadcast
        avg_up = View(avg.reshape(b, 1, 1, 128))
        feat32_sum = self.arm32.run(feat32, avg_up)
        feat32_up = self.conv_head32.run(feat32_sum, in_shift=1)    # nearest x2 folded into the gather
        feat16_sum = self.arm16.run(feat16, feat32_up)
        feat16_up = self.conv_head16.run(feat16_sum, in_shift=1, out=View(fcat, 128, 128))
        return feat8, feat16_up, feat32_up, fcat


class FeatureFusionModule(nn.Module):
    def __init__(self, in_chan, out_chan, *args, **kwargs):
        super().__init__()
        self.convblk = ConvBNReLU(in_chan, out_chan, ks=1, stride=1, padding=0)
        self.conv1 = nn.Conv2d(out_chan, out_chan // 4, kernel_size=1, stride=1, padding=0, bias=False)
        self.conv2 = nn.Conv2d(out_chan // 4, out_chan, kernel_size=1, stride=1, padding=0, bias=False)
        nn.init.kaiming_normal_(self.conv1.weight, a=1)
        nn.init.kaiming_normal_(self.conv2.weight, a=1)

    def run(self, fcat: View) -> View:
        feat = self.convblk.run(fcat)
        c = feat.c
        mean, _ = L.chan_stats(feat.t, c, want_rstd=False)
        a = L.vec_fc(mean, self.conv1.weight.detach().reshape(c // 4, c).contiguous(), None, None, L.ACT_RELU)
        a = L.vec_fc(a, self.conv2.weight.detach().reshape(c, c // 4).contiguous(), None, None, L.ACT_SIGMOID)
        return View(L.residual_combine(feat.t, c, gate=a, gate_plus_one=True))      # feat*atten + feat


class BiSeNet(nn.Module):
    def __init__(self, n_classes, *args, **kwargs):
        super().__init__()
        self.n_classes = n_classes
        self.cp = ContextPath()
        self.ffm = FeatureFusionModule(256, 256)
        self.conv_out = BiSeNetOutput(256, 256, n_classes)
        self.conv_out16 = BiSeNetOutput(128, 64, n_classes)
        self.conv_out32 = BiSeNetOutput(128, 64, n_classes)
        E.install_pack_invalidation(self)

    def run_main(self, x: View):
        """-> (main-head logits NHWC at 1/8 resolution, feat_cp8 view, feat_cp16 view)."""
        _, feat_cp8, feat_cp16, fcat = self.cp.run(x)
        fuse = self.ffm.run(View(fcat))
        return self.conv_out.run(fuse), feat_cp8, feat_cp16

    @torch.no_grad()
    def forward(self, x):
        h, w = x.shape[2:]
        xin = View(L.nchw_to_nhwc(x.contiguous().float(), RGB_PAD))
        out, cp8, cp16 = self.run_main(xin)
        out16 = self.conv_out16.run(cp8)
        out32 = self.conv_out32.run(cp16)
        up = lambda v: L.resize_bilinear_nhwc_to_nchw(v.t, self.n_classes, h, w, align_corners=True)
        return up(out), up(out16), up(out32)

    @torch.no_grad()
    def labels(self, x_nhwc: torch.Tensor, out_hw, lut=None) -> torch.Tensor:
        """Fast path used by FaceParser: normalised NHWC input -> u8 label map [B,H,W]; the full-resolution
        logits are never materialised and the two auxiliary heads (discarded by FaceParser, face_parsing_demo.py:169)
        are skipped."""
        out, _, _ = self.run_main(View(x_nhwc))
        return L.upsample_argmax(out.t, self.n_classes, out_hw[0], out_hw[1], lut)


def bisenet_state_shapes(n_classes=19):
    with torch.device("meta"):
        m = BiSeNet(n_classes)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}

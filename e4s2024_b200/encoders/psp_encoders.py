"""Regional style encoder -- drop-in for `FSEncoder_PSP` (models/encoders/psp_encoders.py:319-401).
Same constructor, forward(x, segmap) -> (codes_vector [B,K,1280], structure_feats zeros [B,512,H/16,W/16])
and state_dict keys (input_layer.{0,2}.weight, body.N.res_layer.*, body.N.shortcut_layer.0.weight)."""
import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..engine import View
from .helpers import PackedConv2d, bottleneck_IR_SE_Ours, get_block

RGB_PAD = 8          # the engine wants cin % 8 == 0: RGB is zero-padded to 8 channels


class FSEncoder_PSP(nn.Module):
    def __init__(self, mode="ir_se", opts=None):
        super().__init__()
        assert mode in ["ir", "ir_se"], "mode should be ir or ir_se"
        if mode != "ir_se":
            raise NotImplementedError("only the 'ir_se' unit (bottleneck_IR_SE_Ours) is on the E4S hot path")
        blocks = [get_block(in_channel=64, depth=128, num_units=3), get_block(in_channel=128, depth=256, num_units=4),
                  get_block(in_channel=256, depth=512, num_units=14), get_block(in_channel=512, depth=512, num_units=3)]
        self.n_styles = 11
        self.input_layer = nn.Sequential(PackedConv2d(3, 64, (3, 3), 1, 1, bias=False), nn.InstanceNorm2d(64), nn.PReLU(64))
        modules = []
        for block in blocks:
            for bottleneck in block:
                modules.append(bottleneck_IR_SE_Ours(bottleneck.in_channel, bottleneck.depth, bottleneck.stride))
        self.body = nn.Sequential(*modules)
        E.install_pack_invalidation(self)

    def get_per_comp_styleCode(self, style_feats, segmap):
        """psp_encoders.py:355-375 on NCHW feats: per-(sample, region) masked mean, zeros for empty regions."""
        f = L.nchw_to_nhwc(style_feats.contiguous().float())
        seg = segmap.contiguous().float()
        codes = torch.zeros(f.shape[0], seg.shape[1], f.shape[3], device=f.device, dtype=torch.float32)
        L.masked_mean(f, f.shape[3], seg, codes, 0)
        return codes

    def run(self, x: View, segmap: torch.Tensor):
        """x: NHWC view with RGB_PAD channels."""
        b = x.bhw[0]
        c0 = E.conv(x, self.input_layer[0].packed(cin_pad=RGB_PAD))
        st = L.chan_stats(c0.t, 64)
        cur = View(L.residual_combine(c0.t, 64, a_stats=st, prelu=self.input_layer[2].weight.detach()))
        seg = segmap.contiguous().float()
        k = seg.shape[1]
        codes = torch.empty(b, k, 256 + 512 + 512, device=x.t.device, dtype=torch.float32)
        off = {6: 0, 20: 256, 23: 768}
        # the three poolings share the mask: reduce it once to a region-membership bit map (K <= 32), one word per pixel
        bits = L.mask_member_bits(seg) if k <= 32 else None
        for i, unit in enumerate(self.body):
            cur = unit.run(cur)
            if i in off:
                if bits is not None:
                    L.masked_mean_bits(cur.t, cur.c, bits, k, codes, off[i])
                else:
                    L.masked_mean(cur.t, cur.c, seg, codes, off[i])
        _, h, w = cur.bhw
        structure_feats = torch.zeros(b, cur.c, h, w, device=x.t.device, dtype=torch.float32)
        return codes, structure_feats

    @torch.no_grad()
    def forward(self, x, segmap):
        xin = View(L.nchw_to_nhwc(x.contiguous().float(), RGB_PAD))
        return self.run(xin, segmap)


def encoder_state_shapes():
    with torch.device("meta"):
        m = FSEncoder_PSP("ir_se", None)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}

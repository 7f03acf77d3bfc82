"""Experiment: is the tensor-core fp32 accumulation biased (round toward zero)?  Least-squares scale of the tcgen05 result
against torch fp64 for several K, and the error left after removing that scale."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from e4s2024_b200 import _lib as L, engine as E
torch.manual_seed(0)
for (cin, cout, hw) in ((64, 64, 64), (128, 128, 64), (256, 256, 32), (512, 512, 32), (512, 512, 16)):
    x = torch.randn(1, hw, hw, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * (1.0 / (9 * cin)) ** 0.5
    pw = E.pack_conv_weight(w)
    y = E.conv(E.View(x), pw, engine="tc").t.permute(0, 3, 1, 2).double()
    y32 = E.conv(E.View(x), pw, engine="f32").t.permute(0, 3, 1, 2).double()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1)
    r = float((y * ref).sum() / (ref * ref).sum()) - 1.0
    r32 = float((y32 * ref).sum() / (ref * ref).sum()) - 1.0
    e0 = float((y - ref).abs().max()); e1 = float((y / (1 + r) - ref).abs().max())
    steps = 9 * cin // 16 * 3
    print(f"K={9*cin:5d} ({steps} accumulate steps): scale bias tc {r:+.3e} (f32 engine {r32:+.3e}); -steps/2*2^-24 = {-steps/2*2**-24:+.3e}; "
          f"max err {e0:.3e} -> {e1:.3e} after unbiasing (scale {float(ref.abs().max()):.2f})")

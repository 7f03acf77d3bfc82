"""ncu aid: one masked up-convolution (hin x hin x cin -> cout, B = 16, face masks) through conv_tc_upz, profiled range only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth, _lib as L, engine as E

hin, cin, cout = (int(v) for v in sys.argv[1:4])
B, K = 16, 12
fir = torch.tensor([1., 3., 3., 1.])
fir = (torch.outer(fir, fir) / 64 * 4).cuda()
labels = synth.make_labels("face", B, K, 512, seed=1)[:, 0].cuda().to(torch.uint8).contiguous()
x = synth.randn("upz.x", (B, hin, hin, cin), 3).cuda()
w = synth.randn("upz.w", (cout, cin, 3, 3), 4, (1.0 / (cin * 9)) ** 0.5).cuda()
smod = (1.0 + 0.3 * synth.randn("upz.s", (B, K, cin), 5)).cuda().contiguous()
demod = (1.0 + 0.2 * synth.randn("upz.d", (B, K, cout), 6)).cuda().contiguous()
noise = synth.randn("upz.n", (1, 1, 2 * hin, 2 * hin), 7).cuda()
nw = torch.tensor([0.1], device="cuda")
bias = synth.randn("upz.b", (cout,), 8, 0.1).cuda()
pw = E.pack_up_weight(w, fir)
kw = dict(up2=True, smod=smod, demod=demod, labels=labels, regions=K, noise=noise, noise_w=nw, ch_shift=bias, act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
cells_total = B * (hin + 1) * (hin + 1)
max_rows = E.pad_to(4 * cells_total, 128)
cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
cells, rows = L.upz_build_rows(labels, hin, hin, max_rows, cnt)
uz = E.UpzRows(cells, rows, cnt, int(cnt.item()), max_rows, max_rows, cells_total)
out = E.View(E.new_nhwc(B, 2 * hin, 2 * hin, cout, "cuda"))
for _ in range(2):
    E.conv(E.View(x), pw, out=out, upz=uz, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
E.conv(E.View(x), pw, out=out, upz=uz, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""ncu aid: one plain convolution through e4s_conv_tc (profiled range only).  args: hin cin cout k stride [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth, _lib as L, engine as E

hin, cin, cout, k, stride = (int(v) for v in sys.argv[1:6])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 16
x = synth.randn("c1.x", (B, hin, hin, cin), 3).cuda()
w = synth.randn("c1.w", (cout, cin, k, k), 4, (1.0 / (cin * k * k)) ** 0.5).cuda()
bias = synth.randn("c1.b", (cout,), 8, 0.1).cuda()
pw = E.pack_conv_weight(w)
kw = dict(stride=stride, ch_shift=bias, act=L.ACT_RELU)
out = E.conv(E.View(x), pw, **kw)
for _ in range(2):
    E.conv(E.View(x), pw, out=out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    E.conv(E.View(x), pw, out=out, **kw)
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1) / 10)
torch.cuda.profiler.start()
E.conv(E.View(x), pw, out=out, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""Timing aid: the un-masked up-convolutions of the 1024^2 generator (B = 16) through the poly-phase halo kernel vs the direct conv_transpose
cell GEMM + FIR pass (E4S_UPZ_DIRECT).  Under ncu (E4S_NCU=1) one direct pass per layer is profiled."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth, _lib as L, engine as E

B = 16
fir = torch.tensor([1., 3., 3., 1.])
fir = (torch.outer(fir, fir) / 64 * 4).cuda()


def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for hin, cin, cout in [(256, 128, 64), (512, 64, 32)]:
    x = synth.randn("upd.x", (B, hin, hin, cin), 3).cuda()
    w = synth.randn("upd.w", (cout, cin, 3, 3), 4, (1.0 / (cin * 9)) ** 0.5).cuda()
    smod = (1.0 + 0.3 * synth.randn("upd.s", (B, 1, cin), 5)).cuda().contiguous()
    demod = (1.0 + 0.2 * synth.randn("upd.d", (B, 1, cout), 6)).cuda().contiguous()
    noise = synth.randn("upd.n", (1, 1, 2 * hin, 2 * hin), 7).cuda()
    nw = torch.tensor([0.1], device="cuda")
    bias = synth.randn("upd.b", (cout,), 8, 0.1).cuda()
    pw = E.pack_up_weight(w, fir)
    kw = dict(up2=True, smod=smod, demod=demod, regions=1, noise=noise, noise_w=nw, ch_shift=bias, act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
    out = E.View(E.new_nhwc(B, 2 * hin, 2 * hin, cout, "cuda"))
    os.environ["E4S_UPZ_DIRECT"] = "0"
    ms_halo = t(lambda: E.conv(E.View(x), pw, out=out, **kw))
    y0 = out.t.clone()
    os.environ["E4S_UPZ_DIRECT"] = "1"
    ms_direct = t(lambda: E.conv(E.View(x), pw, out=out, **kw))
    d = float((out.t - y0).abs().max())
    print(json.dumps({"hin": hin, "cin": cin, "cout": cout, "ms_polyphase_halo": round(ms_halo, 4), "ms_direct": round(ms_direct, 4), "maxdiff": d}), flush=True)
    if os.environ.get("E4S_NCU"):
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        E.conv(E.View(x), pw, out=out, **kw)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

"""Profiling aid: timeline of CTA 0 of the halo kernel on the last generator layer shape (32->32 @1024^2, B=16)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from e4s2024_b200 import _lib as L, engine as E

B, Cin, Cout = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 32
up2 = len(sys.argv) > 3 and sys.argv[3] == "up"
RES = int(sys.argv[4]) if len(sys.argv) > 4 else 1024          # output resolution
H = W = RES // 2 if up2 else RES
x = torch.randn(B, H, W, Cin, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
fir = torch.tensor([1., 3., 3., 1.]); fir = (torch.outer(fir, fir) / 16).cuda()
pw = E.pack_up_weight(w, fir) if up2 else E.pack_conv_weight(w)
smod = torch.rand(B, 1, Cin, device="cuda") + 0.5
demod = torch.rand(B, 1, Cout, device="cuda") + 0.5
noise = torch.randn(1, 1, 2 * H if up2 else H, 2 * W if up2 else W, device="cuda")
nw = torch.tensor([0.1], device="cuda"); bias = torch.randn(Cout, device="cuda")
kw = dict(up2=up2, smod=smod, demod=demod, regions=1, noise=noise, noise_w=nw, ch_shift=bias, act=L.ACT_LRELU, slope=0.2, gain=1.4)
out = E.conv(E.View(x), pw, engine="tc", **kw)
torch.cuda.synchronize()
cap = 64
LAPS = {0: ("producer", {0: "loop", 1: "wait hempty", 2: "convert+st.shared", 3: "fence+arrive", 4: "prefetch issue"}),
        1: ("epilogue", {0: "job setup", 3: "sv rebuild", 1: "wait afull", 2: "tmem ld + math + stores + arrive"}),
        2: ("mma", {0: "loop", 1: "wait aempty", 2: "wait hfull", 4: "wait bfull", 5: "issue (chunks)", 3: "issue (last chunk) + commits"}),
        3: ("loader", {0: "wait bempty", 1: "issue"})}
for flags in [int(f) for f in os.environ.get("TRACE_FLAGS", "0").split(",")]:
    L.lib().e4s_debug_halo_flags(flags)
    buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
    L.lib().e4s_debug_halo_trace(C.c_void_p(buf.data_ptr()), cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E.conv(E.View(x), pw, engine="tc", out=out, **kw); e1.record()
    torch.cuda.synchronize()
    L.lib().e4s_debug_halo_trace(None, 0)
    L.lib().e4s_debug_halo_flags(0)
    v = buf.cpu().numpy()
    jobs = max(int(v[32]), 1)
    print(f"=== flags {flags}: kernel {e0.elapsed_time(e1):.3f} ms with accounting, CTA 0 ran {jobs} jobs; cycles per job:")
    for r, (name, laps) in LAPS.items():
        tot = sum(v[r * 8 + k] for k in laps) / jobs
        print(f"  {name:9s} total {tot:8.0f} | " + " | ".join(f"{n} {v[r * 8 + k] / jobs:.0f}" for k, n in laps.items()))

"""Profiling aid: timeline of CTA 0 of the halo kernel on the last generator layer shape (32->32 @1024^2, B=16)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from e4s2024_b200 import _lib as L, engine as E

B, H, W, Cin, Cout = 16, 1024, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 32
up2 = len(sys.argv) > 3 and sys.argv[3] == "up"
if up2:
    H = W = 512
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
for flags in (0, 1, 8, 16, 24, 7):
    L.lib().e4s_debug_halo_flags(flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E.conv(E.View(x), pw, engine="tc", out=out, **kw); e1.record(); torch.cuda.synchronize()
    print(f"flags {flags}: kernel {e0.elapsed_time(e1):.3f} ms (no trace)")
L.lib().e4s_debug_halo_flags(0)
cap = 40000
buf = torch.zeros(2 * cap, dtype=torch.int64, device="cuda")
L.lib().e4s_debug_halo_trace(C.c_void_p(buf.data_ptr()), cap)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); E.conv(E.View(x), pw, engine="tc", out=out, **kw); e1.record()
torch.cuda.synchronize()
L.lib().e4s_debug_halo_trace(None, 0)
print("kernel ms", e0.elapsed_time(e1))
rec = buf.cpu().numpy().reshape(-1, 2)
rec = rec[rec[:, 1] != 0]
role = rec[:, 0] >> 48; it = (rec[:, 0] >> 16) & 0xffffff; ev = rec[:, 0] & 0xffff; t = rec[:, 1] - rec[:, 1].min()
names = {0: "producer", 1: "epilogue", 2: "mma"}
for r in (0, 1, 2):
    m = role == r
    print(f"--- {names[r]}: {m.sum()} records")
    evs = sorted(set(ev[m]))
    # mean duration between consecutive events within a job, and job period
    per = {}
    for e in evs:
        tt = t[m & (ev == e)]; ii = it[m & (ev == e)]
        order = np.argsort(ii); per[e] = (ii[order], tt[order])
    base = per[evs[0]]
    n = min(len(v[0]) for v in per.values())
    ref = per[evs[0]][1][:n]
    for e in evs[1:]:
        d = per[e][1][:n] - ref
        print(f"  ev{evs[0]}->ev{e}: mean {d[5:].mean():9.0f} cyc  (p50 {np.median(d[5:]):9.0f})")
    period = np.diff(base[1][:n])
    print(f"  job period: mean {period[5:].mean():9.0f} cyc over {n} jobs")

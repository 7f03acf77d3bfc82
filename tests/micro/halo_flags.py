"""Profiling aid: halo kernel time on one layer shape under the debug flags (which role bounds the job period?).
   bit0 skip epilogue math+stores, bit1 skip tcgen05.ld, bit2 skip MMAs, bit3 skip stores, bit4 skip math,
   bit5 producers skip loads+convert, bit6 weight loader skips the bulk copies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import _lib as L, engine as E

def run(Cin, Cout, res, up2, B=16):
    H = W = res // 2 if up2 else res
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    fir = torch.tensor([1., 3., 3., 1.]); fir = (torch.outer(fir, fir) / 16).cuda()
    pw = E.pack_up_weight(w, fir) if up2 else E.pack_conv_weight(w)
    smod = torch.rand(B, 1, Cin, device="cuda") + 0.5
    demod = torch.rand(B, 1, Cout, device="cuda") + 0.5
    noise = torch.randn(1, 1, res, res, device="cuda")
    nw = torch.tensor([0.1], device="cuda"); bias = torch.randn(Cout, device="cuda")
    kw = dict(up2=up2, smod=smod, demod=demod, regions=1, noise=noise, noise_w=nw, ch_shift=bias, act=L.ACT_LRELU, slope=0.2, gain=1.4)
    out = E.conv(E.View(x), pw, engine="tc", **kw)
    torch.cuda.synchronize()
    res_ms = {}
    for flags in (0, 1, 7, 32, 64, 96, 97, 103, 33, 65):
        L.lib().e4s_debug_halo_flags(flags)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); E.conv(E.View(x), pw, engine="tc", out=out, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res_ms[flags] = min(ts)
    L.lib().e4s_debug_halo_flags(0)
    print(f"{Cin}->{Cout} @{res} up={up2}: " + "  ".join(f"f{k}={v:.3f}" for k, v in res_ms.items()), flush=True)

for a in sys.argv[1:]:
    cin, cout, res, up = a.split(",")
    run(int(cin), int(cout), int(res), up == "up")

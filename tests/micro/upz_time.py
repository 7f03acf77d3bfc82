"""Timing aid: the masked up-convolutions of the 1024^2 generator (B = 16) as the poly-phase kernels vs the (cell, region) conv_transpose
GEMM + FIR pass (csrc/conv_tc_upz.cu), per mask family.  Under ncu (E4S_NCU=1, --profile-from-start off) one upz pass per layer is
profiled so that the GEMM and the FIR pass can be read separately."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth, _lib as L, engine as E

B, K = 16, 12
LAYERS = [(8, 512, 512), (16, 512, 512), (32, 512, 512), (64, 512, 256), (128, 256, 128)]      # (hin, cin, cout)
kinds = sys.argv[1:] or ["face", "blocky"]
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


for kind in kinds:
    labels = synth.make_labels(kind, B, K, 512, seed=1)[:, 0].cuda().to(torch.uint8).contiguous()
    for hin, cin, cout in LAYERS:
        x = synth.randn("upz.x", (B, hin, hin, cin), 3).cuda()
        w = synth.randn("upz.w", (cout, cin, 3, 3), 4, (1.0 / (cin * 9)) ** 0.5).cuda()
        smod = (1.0 + 0.3 * synth.randn("upz.s", (B, K, cin), 5)).cuda().contiguous()
        demod = (1.0 + 0.2 * synth.randn("upz.d", (B, K, cout), 6)).cuda().contiguous()
        noise = synth.randn("upz.n", (1, 1, 2 * hin, 2 * hin), 7).cuda()
        nw = torch.tensor([0.1], device="cuda")
        bias = synth.randn("upz.b", (cout,), 8, 0.1).cuda()
        pw = E.pack_up_weight(w, fir)
        kw = dict(up2=True, smod=smod, demod=demod, labels=labels, regions=K, noise=noise, noise_w=nw, ch_shift=bias,
                  act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
        cells_total = B * (hin + 1) * (hin + 1)
        max_rows = E.pad_to(13 * cells_total, 128)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        cells, rows = L.upz_build_rows(labels, hin, hin, max_rows, cnt)
        uz = E.UpzRows(cells, rows, cnt, int(cnt.item()), max_rows, max_rows, cells_total)
        out = E.View(E.new_nhwc(B, 2 * hin, 2 * hin, cout, "cuda"))
        ms_old = t(lambda: E.conv(E.View(x), pw, out=out, **kw))
        ms_new = t(lambda: E.conv(E.View(x), pw, out=out, upz=uz, **kw))
        def build():
            cnt.zero_()
            L.upz_build_rows(labels, hin, hin, max_rows, cnt)
        ms_build = t(build)
        print(json.dumps({"mask": kind, "hin": hin, "cin": cin, "cout": cout, "rows_per_cell": round(uz.count / cells_total, 3),
                          "ms_polyphase": round(ms_old, 4), "ms_upz": round(ms_new, 4), "ms_build_rows": round(ms_build, 4)}), flush=True)
        if os.environ.get("E4S_NCU"):
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            E.conv(E.View(x), pw, out=out, upz=uz, **kw)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()

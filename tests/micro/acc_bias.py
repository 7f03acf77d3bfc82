"""Experiment: the tensor core's fp32 accumulate truncates (rounds toward zero).  For both operand formats of the 3-pass split
(bf16 hi/lo, fp16 hi/lo), several K and several operand distributions: least-squares scale of the RAW tcgen05 result
(E4SConv.tc_unbias < 0) against torch fp64 -> bias per accumulate step, and the max error with and without the library's
per-format correction.  Writes gpurun_out/acc_bias.json.

    python tests/micro/acc_bias.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from e4s2024_b200 import _lib as L, engine as E

torch.manual_seed(0)
SHAPES = ((32, 32, 64), (64, 64, 64), (128, 128, 64), (256, 256, 32), (512, 512, 32), (512, 512, 16))


def operands(dist, cin, cout, hw):
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + hw)
    x = torch.randn(1, hw, hw, cin, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (1.0 / (9 * cin)) ** 0.5
    if dist == "normal":
        pass
    elif dist == "post_lrelu":            # non-zero-mean activations: what a StyledConv hands the next layer
        x = F.leaky_relu(x, 0.2) * 2 ** 0.5
    elif dist == "post_relu_bn":          # BiSeNet: ReLU output, weights with a common sign bias
        x = F.relu(x + 0.5)
        w = w + 0.3 * w.abs().mean()
    elif dist == "heavy_tail_w":          # student-t(3) weights
        t = torch.distributions.StudentT(3.0).sample(w.shape).to("cuda")
        w = t * (1.0 / (9 * cin)) ** 0.5 / 1.7
    elif dist == "small_x":               # activations far below 1 (fp16 lo parts go subnormal)
        x = x * 1e-3
    return x.contiguous(), w.contiguous()


def run(fmt, dist, cin, cout, hw):
    x, w = operands(dist, cin, cout, hw)
    pw = E.pack_conv_weight(w, tc_fmt=fmt)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1)
    out = {}
    for tag, ov in (("raw", -1.0), ("corrected", None)):
        E.TC_UNBIAS_OVERRIDE = ov
        y = E.conv(E.View(x), pw, engine="tc").t.permute(0, 3, 1, 2).double()
        out[tag] = (float((y * ref).sum() / (ref * ref).sum()) - 1.0, float((y - ref).abs().max()))
    E.TC_UNBIAS_OVERRIDE = None
    y32 = E.conv(E.View(x), pw, engine="f32").t.permute(0, 3, 1, 2).double()
    e32 = float((y32 - ref).abs().max())
    nc = cout <= 64 and cin % 64 == 0 or (cin == 32 and cout == 32)      # halo hi|lo N-merge: 2 accumulate steps per K step
    steps = 9 * cin // 16 * (1 if fmt == L.TC_F16 else (2 if nc else 3))   # accumulate steps into the MAIN accumulator
    return {"fmt": "f16" if fmt == L.TC_F16 else "bf16", "dist": dist, "cin": cin, "cout": cout, "K": 9 * cin, "steps": steps,
            "scale_bias_raw": out["raw"][0], "bias_per_step": out["raw"][0] / steps, "scale_bias_corrected": out["corrected"][0],
            "maxerr_raw": out["raw"][1], "maxerr_corrected": out["corrected"][1], "maxerr_f32_engine": e32,
            "ref_absmax": float(ref.abs().max())}


def main():
    rows = []
    for fmt in (L.TC_BF16, L.TC_F16):
        for dist in ("normal", "post_lrelu", "post_relu_bn", "heavy_tail_w", "small_x"):
            for (cin, cout, hw) in SHAPES:
                r = run(fmt, dist, cin, cout, hw)
                rows.append(r)
                print(f"{r['fmt']:4s} {dist:13s} K={r['K']:5d} steps={r['steps']:4d}: raw scale bias {r['scale_bias_raw']:+.3e} "
                      f"({r['bias_per_step']:+.2e}/step) corrected {r['scale_bias_corrected']:+.3e}; max err raw {r['maxerr_raw']:.2e} "
                      f"corrected {r['maxerr_corrected']:.2e} f32 engine {r['maxerr_f32_engine']:.2e} (|ref| max {r['ref_absmax']:.2f})")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/acc_bias.json", "w"), indent=1)


if __name__ == "__main__":
    main()

"""HBM-bound kernels of the path at the generator's largest shapes: CUDA-event time, algorithmic bytes, GB/s against the measured
copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  Run plain for the numbers, or under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` for the DRAM traffic of the same launches.
Writes gpurun_out/pointwise_bw.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import _lib as L, synth

dev = "cuda"
peak = 6555.2
p = os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
k1 = torch.tensor([1., 3., 3., 1.])
fir = (torch.outer(k1, k1) / 64).to(dev)
res = []


def t(name, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    r = {"kernel": name, "ms": ms, "alg_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_measured_copy_bw": nbytes / ms / 1e6 / peak}
    res.append(r); print(json.dumps(r))


B = 16
# Blur after the last conv_transpose (model.py:300): [16*32, 1025, 1025] -> [.., 1024, 1024]
x = torch.randn(B, 32, 1025, 1025, device=dev)
t("upfirdn2d_tile_kernel<1> blur 1025^2->1024^2 x512 planes", lambda: L.upfirdn2d(x, fir * 4, 1, 1, 1, 1), 4 * (x.numel() + B * 32 * 1024 * 1024))
del x
# Upsample of the skip image (model.py:34-53): [16,3,512,512] -> 1024^2
s = torch.randn(B, 3, 512, 512, device=dev)
t("upfirdn2d_tile_kernel<2> skip 512^2->1024^2 x48 planes", lambda: L.upfirdn2d(s, fir * 4, 2, 1, 2, 1), 4 * (s.numel() + B * 3 * 1024 * 1024))
# FusedLeakyReLU on the largest activation
a = torch.randn(B, 32, 1024, 1024, device=dev)
bias = torch.randn(32, device=dev)
t("bias_act_vec4_kernel 16x32x1024^2", lambda: L.bias_act(a, bias, 0.2, 2 ** 0.5), 8 * a.numel())
# tensor2im / im2tensor
img = torch.randn(B, 3, 1024, 1024, device=dev)
t("tensor2im_kernel 16x3x1024^2", lambda: L.tensor2im_u8(img, True), 5 * img.numel())
u8 = torch.randint(0, 255, (B, 1024, 1024, 3), dtype=torch.uint8, device=dev)
t("im2tensor_kernel 16x1024^2x3 (two outputs)", lambda: L.im2tensor(u8), 9 * u8.numel())
# standalone ToRGB at 256^2 (128 channels in) with skip: reads x once, writes 3 planes
xh = torch.randn(B, 256, 256, 128, device=dev)
smod = torch.randn(B, 128, device=dev); wrgb = torch.randn(3, 128, device=dev); skip = torch.randn(B, 3, 128, 128, device=dev)
rgb = torch.empty(B, 3, 256, 256, device=dev); b3 = torch.zeros(3, device=dev)
t("torgb_kernel 16x256^2x128", lambda: L.torgb(xh, 128, smod, wrgb, None, 1, (0, 0), None, 0, b3, skip, fir * 4, rgb, False), 4 * (xh.numel() + rgb.numel() + skip.numel()))
# masked mean at the encoder's first tap (256 ch @ 64^2, 12 regions, 512^2 mask)
f = torch.randn(B, 64, 64, 256, device=dev)
mask = synth.onehot(synth.make_labels("face", B, 12, 512, seed=1), 12).to(dev)
codes = torch.zeros(B, 12, 1280, device=dev)
t("masked_mean_kernel 16x64^2x256, K=12", lambda: L.masked_mean(f, 256, mask, codes, 0), 4 * f.numel() + 4 * B * 12 * 64 * 64)
# labels -> one-hot, mask -> labels
lab = torch.randint(0, 12, (B, 512, 512), dtype=torch.uint8, device=dev)
t("labels_to_onehot 16x12x512^2", lambda: L.labels_to_onehot(lab, 12), lab.numel() + 4 * 12 * lab.numel())
t("mask_labels 16x12x512^2", lambda: L.mask_labels(mask), 4 * mask.numel() + lab.numel())
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"hbm_gbs_measured": peak, "rows": res}, open("gpurun_out/pointwise_bw.json", "w"), indent=1)

"""Round-2 golden vectors from the REFERENCE's own code, executed in the build container (TEST INFRASTRUCTURE ONLY):
    python oracle/make_golden_r2.py   -> tests/golden/im2tensor.npz
TO_TENSOR / NORMALIZE (datasets/dataset.py:45,52-56 as composed in face_swap_video_pipeline.py:338-346)."""
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
for _m in ("matplotlib", "matplotlib.pyplot"):          # absent in this image and unused here: import shim only
    sys.modules.setdefault(_m, types.ModuleType(_m))

from oracle import e4s_oracle as orc  # noqa: E402
from datasets.dataset import TO_TENSOR, NORMALIZE  # noqa: E402
import torchvision.transforms as transforms  # noqa: E402

g = np.random.default_rng(123)
x = g.integers(0, 256, (2, 16, 24, 3), dtype=np.uint8)
x[0, 0, :8, 0] = [0, 1, 2, 127, 128, 254, 255, 63]      # every byte value is covered by the second fixture below
allb = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
x = np.concatenate([x, np.zeros((1, 16, 24, 3), np.uint8)], 0)
x[2, :, :16] = allb[0]
y01 = torch.stack([TO_TENSOR(Image.fromarray(x[i])) for i in range(x.shape[0])])
yn = torch.stack([transforms.Compose([TO_TENSOR, NORMALIZE])(Image.fromarray(x[i])) for i in range(x.shape[0])])
o01, on = orc.to_tensor_normalize(torch.from_numpy(x))
print("oracle vs reference: x01", float((o01 - y01).abs().max()), "normalised", float((on - yn).abs().max()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "im2tensor.npz"), x=x, y01=y01.numpy(), ynorm=yn.numpy())

# ---- SoftErosion (utils/paste_back_tricks.py:17-42) and the Laplacian-pyramid blend (swap_face_fine/multi_band_blending.py) -------------
import cv2  # noqa: E402
from utils.paste_back_tricks import SoftErosion as RefSoftErosion  # noqa: E402
from swap_face_fine.multi_band_blending import Laplacian_Pyramid_Blending_with_mask as ref_lpb, blending as ref_blending  # noqa: E402
from e4s2024_b200 import synth  # noqa: E402

arr = {}
lab = synth.face_labels(2, 96, seed=5)
fg = ((lab != 0) & (lab != 4)).float()                       # a face-shaped foreground mask [2,1,96,96]
worst = 0.0
for name, kw in (("default", {}), ("k7_it3", dict(kernel_size=7, threshold=0.5, iterations=3))):
    y, mk = RefSoftErosion(**kw)(fg.clone())
    oy, omk = orc.soft_erosion(fg.clone(), **kw)
    worst = max(worst, float((y - oy).abs().max()), float((mk != omk).sum()))
    arr[name + "_y"], arr[name + "_mask"] = y.numpy(), mk.numpy()
    arr[name + "_cfg"] = np.array([kw.get("kernel_size", 15), kw.get("iterations", 1)], np.int64)
    arr[name + "_thr"] = np.float32(kw.get("threshold", 0.6))
print("soft_erosion: oracle vs reference max|diff| (and mask mismatches)", worst)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "soft_erosion.npz"), x=fg.numpy(), **arr)

rng = np.random.default_rng(7)
worst_d = worst_u = 0.0
for shape in ((8, 8, 3), (7, 9, 3), (2, 2, 3), (1, 1, 3), (16, 12), (33, 20, 3)):
    a = (rng.random(shape) * 255).astype(np.float32)
    worst_d = max(worst_d, float(np.abs(orc.pyr_down(a) - cv2.pyrDown(a)).max()))
    worst_u = max(worst_u, float(np.abs(orc.pyr_up(a) - cv2.pyrUp(a)).max()))
    a8 = a.astype(np.uint8)
    assert (orc.pyr_down(a8.astype(np.float32), round_u8=True).astype(np.uint8) != cv2.pyrDown(a8)).sum() == 0
print("pyr_down / pyr_up: oracle vs cv2 max|diff| on [0,255] data", worst_d, worst_u, "(uint8 pyrDown exact)")

S = 64
A8 = (synth.smooth_image_u8("blend.A", 1, S, 21)[0]).numpy()                    # uint8 HxWx3, as np.array(PIL)
Bf = ((synth.smooth_image("blend.B", 1, S, 22)[0].permute(1, 2, 0).numpy() + 1) * 127.5).astype(np.float32)     # float32, like the pipelines' composite
mk = np.repeat(np.clip(synth.smooth_image("blend.m", 1, S, 23)[0, :1].permute(1, 2, 0).numpy() * 2, 0, 1), 3, axis=2).astype(np.float32)
ref = ref_lpb(A8, Bf, mk, 5)
mine = orc.laplacian_pyramid_blend(A8, Bf, mk, 5)
print("laplacian blend (64^2, 5 levels, uint8 A / float B): oracle vs reference max|diff|", float(np.abs(ref - mine).max()))
Af = A8.astype(np.float32)
ref_f = ref_lpb(Af, Bf.astype(np.float32), mk, 5)
print("laplacian blend (float A): oracle vs reference max|diff|", float(np.abs(ref_f - orc.laplacian_pyramid_blend(Af, Bf.astype(np.float32), mk, 5)).max()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "laplacian_blend.npz"), A8=A8, B=Bf.astype(np.float32), m=mk, out_u8A=ref.astype(np.float32),
                    out_fA=ref_f.astype(np.float32), levels=np.array(5))
# the full 1024^2 / 10-level `blending` of the pipelines, oracle vs reference (not stored: 12 MB; the GPU test recomputes the oracle)
A1 = synth.smooth_image_u8("blend.A1", 1, 1024, 24)[0].numpy()
B1 = ((synth.smooth_image("blend.B1", 1, 1024, 25)[0].permute(1, 2, 0).numpy() + 1) * 127.5).astype(np.float32)
m1 = np.repeat(np.clip(synth.smooth_image("blend.m1", 1, 1024, 26)[0, :1].permute(1, 2, 0).numpy() * 2, 0, 1), 3, axis=2).astype(np.float32)
r1, o1 = ref_blending(A1, B1, m1), orc.blending(A1, B1, m1)
print("blending 1024^2: uint8 pixels differing", int((r1 != o1).sum()), "of", r1.size, "max |diff|", int(np.abs(r1.astype(int) - o1.astype(int)).max()))

"""Golden vectors for swap_comp_style_vector from the REFERENCE's own function (swap_face_fine/swap_face_mask.py:336-367),
executed in the build container:  python oracle/make_golden_swap.py   -> tests/golden/swap_comp_style_vector.npz
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from e4s2024_b200 import synth  # noqa: E402
from oracle import e4s_oracle as orc  # noqa: E402
from swap_face_fine.swap_face_mask import swap_comp_style_vector as ref_fn  # noqa: E402

CASES = [([1, 2, 3, 5, 6, 9], False, False), ([1, 2, 3, 5, 6, 9], True, False), ([4, 8, 9, 10], False, True), ([], False, True)]
arrs, worst = {}, 0.0
for i, (comps, below, empty_mouth) in enumerate(CASES):
    a = synth.randn(f"swapsv.t{i}", (1, 12, 512), 40 + i)
    b = synth.randn(f"swapsv.s{i}", (1, 12, 512), 50 + i)
    if empty_mouth:
        b[:, 9, :] = 0
    y = ref_fn(a, b, comps, belowFace_interpolation=below)
    worst = max(worst, float((y - orc.swap_comp_style_vector(a, b, comps, below)).abs().max()))
    arrs[f"t{i}"], arrs[f"s{i}"], arrs[f"y{i}"] = a.numpy(), b.numpy(), y.numpy()
    arrs[f"cfg{i}"] = np.array([int(below)] + list(comps), dtype=np.int64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "swap_comp_style_vector.npz"), n=np.array(len(CASES)), **arrs)
print("oracle vs reference: max|diff|", worst)

# ---- tensor2im (utils/torch_utils.py:64-76): values around the clamp edges and the rounding boundaries ----------------------
import types  # noqa: E402
for _m in ("matplotlib", "matplotlib.pyplot"):          # absent in this image and unused by tensor2im: import shim only
    sys.modules.setdefault(_m, types.ModuleType(_m))
from utils.torch_utils import tensor2im as ref_tensor2im  # noqa: E402

x = synth.randn("tensor2im.x", (2, 3, 8, 12), 60) * 0.8
x[0, 0, 0, :6] = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, 254.5 / 127.5 - 1])
ys = np.stack([np.array(ref_tensor2im(x[i])) for i in range(x.shape[0])])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tensor2im.npz"), x=x.numpy(), y=ys)
print("tensor2im golden", ys.shape, ys.dtype, "oracle mismatches:", int((orc.tensor2im_u8(x).numpy() != ys).sum()))

# ---- morphology (utils/morphology.py): flat / holed / non-flat elements, even sizes, both engines, constant border ----------
from utils.morphology import dilation as ref_dil, erosion as ref_ero  # noqa: E402

xm = (synth.randn("morph.x", (2, 2, 19, 23), 70) > 0.3).float() * synth.randn("morph.v", (2, 2, 19, 23), 71).abs()
cases_m = {"ones5": dict(kernel=torch.ones(5, 5)),
           "cross3": dict(kernel=torch.tensor([[0., 1, 0], [1, 1, 1], [0, 1, 0]])),
           "even4x6": dict(kernel=torch.ones(4, 6)),
           "nonflat": dict(kernel=torch.ones(3, 3), structuring_element=torch.tensor([[0., 0.1, 0], [0.1, 0.3, 0.1], [0, 0.1, 0]])),
           "const": dict(kernel=torch.ones(3, 5), border_type="constant", border_value=0.5),
           "origin": dict(kernel=torch.ones(3, 3), origin=[0, 2])}
arrm, worst = {"x": xm.numpy()}, 0.0
for name, kw in cases_m.items():
    for eng in ("unfold", "convolution"):
        d, e = ref_dil(xm, engine=eng, **kw), ref_ero(xm, engine=eng, **kw)
        okw = {k: v for k, v in kw.items() if k != "kernel"}
        worst = max(worst, float((d - orc.morphology(xm, kw["kernel"], True, **okw)).abs().max()),
                    float((e - orc.morphology(xm, kw["kernel"], False, **okw)).abs().max()))
        if eng == "unfold":
            arrm[name + "_dil"], arrm[name + "_ero"] = d.numpy(), e.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "morphology.npz"), **arrm)
print("morphology: oracle vs reference (both engines) max|diff|", worst)

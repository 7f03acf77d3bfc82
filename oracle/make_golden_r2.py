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

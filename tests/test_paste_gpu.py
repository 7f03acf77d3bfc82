"""GPU, full size: the pipelines' `blending` (1024^2, 10 pyramid levels, uint8 target + float composite + float border mask) and
SoftErosion on a 1024^2 mask vs the oracle (pinned on the reference's functions / cv2 by oracle/make_golden_r2.py)."""
import numpy as np
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc

pytestmark = pytest.mark.gpu


def test_blending_1024_vs_oracle():
    from e4s2024_b200.multi_band_blending import blending
    A = synth.smooth_image_u8("blend.A1", 2, 1024, 24)                                    # [2,1024,1024,3] uint8
    B = ((synth.smooth_image("blend.B1", 2, 1024, 25) + 1) * 127.5).float()               # [2,3,1024,1024]
    m = (synth.smooth_image("blend.m1", 2, 1024, 26)[:, :1] * 2).clamp(0, 1).repeat(1, 3, 1, 1)
    out = blending(A.permute(0, 3, 1, 2).contiguous().cuda(), B.cuda(), m.cuda())
    assert out.dtype == torch.uint8 and out.shape == (2, 3, 1024, 1024)
    ref = orc.blending(A[0].numpy(), B[0].permute(1, 2, 0).numpy(), m[0].permute(1, 2, 0).numpy())
    d = (out[0].permute(1, 2, 0).cpu().numpy().astype(int) - ref.astype(int))
    print(f"blending 1024^2: {int((d != 0).sum())} of {d.size} uint8 values differ, max |diff| {int(np.abs(d).max())}")
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 1e-4                                 # truncation of values on an integer boundary
    solo = blending(A[1:].permute(0, 3, 1, 2).contiguous().cuda(), B[1:].cuda(), m[1:].cuda())
    assert torch.equal(solo[0], out[1])


def test_soft_erosion_1024_vs_oracle():
    from e4s2024_b200.utils.paste_back_tricks import SoftErosion
    lab = synth.face_labels(2, 1024, seed=8)
    fg = ((lab != 0) & (lab != 4)).float()
    y, mk = SoftErosion().cuda()(fg.cuda())
    ry, rmk = orc.soft_erosion(fg)
    bad = mk.cpu() != rmk
    assert int(bad.sum()) <= 8
    assert float((y.cpu() - ry).abs()[~bad].max()) < 1e-5

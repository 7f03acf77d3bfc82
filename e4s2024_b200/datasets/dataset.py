"""Label-space helpers on either side of the hot path (reference: datasets/dataset.py:58-108).
The 19-class face-parsing ids are folded into the 12 E4S regions with a 256-entry LUT, which the
parser applies on the GPU inside the upsample+argmax kernel; the numpy form below keeps the
reference's call signature for host-side callers."""
import numpy as np

# 19-class id -> 12-region id (dataset.py:68-106): lip, eyebrows, eyes, hair, nose, skin, ears, neck, teeth, glasses, earrings
SEG19_TO_SEG12 = np.zeros(256, dtype=np.uint8)
for _dst, _srcs in enumerate(((0,), (12, 13), (2, 3), (4, 5), (17,), (10,), (1,), (7, 8), (14,), (11,), (6,), (9,))):
    for _s in _srcs:
        SEG19_TO_SEG12[_s] = _dst


def ffhq_masks_to_faceParser_mask_detailed(mask):
    """mask: integer array [H,W] of face-parsing ids -> same shape, 12-region ids."""
    return SEG19_TO_SEG12[np.asarray(mask)].astype(np.asarray(mask).dtype, copy=False)


# the reference spells the name with two leading underscores
globals()["__ffhq_masks_to_faceParser_mask_detailed"] = ffhq_masks_to_faceParser_mask_detailed

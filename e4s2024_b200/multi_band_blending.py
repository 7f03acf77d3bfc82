"""Laplacian-pyramid (multi-band) blending of the swapped face into the target frame -- drop-in for
`swap_face_fine/multi_band_blending.py` (Laplacian_Pyramid_Blending_with_mask :6-49, blending :52-74), batched and on the GPU.

The reference works on numpy HxWx3 arrays through cv2.pyrDown / cv2.pyrUp; here images are CUDA tensors [B,C,H,W] and every pyramid
level is one kernel with cv2's arithmetic (e4s_pyr_down_f32 / e4s_pyr_up_f32 / e4s_pyr_blend_f32: [1 4 6 4 1] Gaussian,
BORDER_REFLECT_101, the Laplacian subtraction and the reconstruction add fused into the up-sampling).  uint8 inputs keep cv2's
per-level integer rounding of their Gaussian pyramid (the pipelines pass `np.array(PIL image)` for the target frame)."""
import torch

from . import _lib as L


def _as_planes(x: torch.Tensor):
    """-> (fp32 [B,C,H,W] contiguous, was_uint8)."""
    if x.dim() == 3:
        x = x[None]
    u8 = x.dtype == torch.uint8
    return x.float().contiguous(), u8


def _gauss_pyramid(x: torch.Tensor, levels: int, round_u8: bool):
    out = [x]
    for _ in range(levels):
        x = L.pyr_down(x, round_u8)
        out.append(x)
    return out


@torch.no_grad()
def Laplacian_Pyramid_Blending_with_mask(A: torch.Tensor, B: torch.Tensor, m: torch.Tensor, num_levels: int = 6) -> torch.Tensor:
    """A, B [B,C,H,W] (uint8 or float, values in [0,255]), m [B,C,H,W] or [B,1,H,W] float in [0,1] -> blended fp32 [B,C,H,W]."""
    A, a8 = _as_planes(A)
    B, b8 = _as_planes(B)
    m, _ = _as_planes(m)
    if A.shape != B.shape or m.shape[0] != A.shape[0] or m.shape[2:] != A.shape[2:] or m.shape[1] not in (1, A.shape[1]):
        raise L.E4SError(f"blend: shapes {tuple(A.shape)}, {tuple(B.shape)}, {tuple(m.shape)} do not match")
    if num_levels < 1:
        raise L.E4SError("blend: num_levels must be >= 1")
    gpA, gpB, gpM = _gauss_pyramid(A, num_levels, a8), _gauss_pyramid(B, num_levels, b8), _gauss_pyramid(m, num_levels, False)
    # Laplacian levels (coarse -> fine), blended per level, then reconstructed bottom-up (:22-47)
    ls = L.pyr_blend(gpA[num_levels - 1], gpB[num_levels - 1], gpM[num_levels - 1])
    for i in range(num_levels - 1, 0, -1):
        la = L.pyr_up(gpA[i], gpA[i - 1], mode=1)
        lb = L.pyr_up(gpB[i], gpB[i - 1], mode=1)
        ls = L.pyr_up(ls, L.pyr_blend(la, lb, gpM[i - 1]), mode=2)
    return ls


@torch.no_grad()
def blending(full_img: torch.Tensor, ori_img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """multi_band_blending.py:52-74 for the 1024x1024 crops the pipelines blend (their cv2.resize calls to and from 1024^2 are identities
    there): 10 pyramid levels, clip to [0,255], uint8 (truncation) -> uint8 [B,C,1024,1024]."""
    if tuple(full_img.shape[-2:]) != (1024, 1024) or tuple(ori_img.shape[-2:]) != (1024, 1024):
        raise L.E4SError("blending: the GPU path takes the pipelines' 1024x1024 crops (resize outside for other sizes)")
    img = Laplacian_Pyramid_Blending_with_mask(full_img, ori_img, mask.float(), 10)
    return img.clamp_(0, 255).to(torch.uint8)

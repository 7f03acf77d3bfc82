"""FaceParser front-end -- drop-in for `swap_face_fine/face_parsing/face_parsing_demo.py`
(BicubicDownSample :15-84, FaceParser :127-176, init_faceParsing_pretrained_model :180-185,
faceParsing_demo :187-200), plus a batched tensor entry point (`FaceParser.parse_batch`) that the
reference lacks (it is PIL, batch 1): bicubic 1024->512 + clamp + normalise, BiSeNet, x8
align_corners upsample + argmax + seg19->seg12 LUT, all on the GPU."""
import math

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from ..datasets.dataset import SEG19_TO_SEG12, ffhq_masks_to_faceParser_mask_detailed
from .model import BiSeNet, seg_mean, seg_std
from .resnet import RGB_PAD


def bicubic_taps(factor, a=-0.5):
    """The 4*factor taps of face_parsing_demo.py:16-35 (cubic kernel sampled at (i - size/2 + 0.5)/factor)."""
    size = factor * 4
    ks = []
    for i in range(size):
        t = abs((i - math.floor(size / 2) + 0.5) / factor)
        if t <= 1.0:
            v = (a + 2.0) * t ** 3 - (a + 3.0) * t ** 2 + 1
        elif t < 2.0:
            v = a * t ** 3 - 5.0 * a * t ** 2 + 8.0 * a * t - 4.0 * a
        else:
            v = 0.0
        ks.append(v)
    k = torch.tensor(ks, dtype=torch.float32)
    return k / torch.sum(k)


class BicubicDownSample(nn.Module):
    def __init__(self, factor=4, cuda=True, padding="reflect"):
        super().__init__()
        if padding != "reflect":
            raise NotImplementedError("only reflect padding is used by the parser")
        self.factor = factor
        self.taps = bicubic_taps(factor)
        self.cuda = ".cuda" if cuda else ""
        self.padding = padding

    def forward(self, x, nhwc=False, clip_round=False, byte_output=False):
        """x NCHW [B,3,H,W] float on CUDA -> [B,3,H/f,W/f] (face_parsing_demo.py:46-84, plain path)."""
        if nhwc or clip_round or byte_output:
            raise NotImplementedError("nhwc / clip_round / byte_output are unused on the hot path")
        dev = x.device
        zero, one = torch.zeros(3, device=dev), torch.ones(3, device=dev)
        y = L.bicubic_down_norm(x.contiguous().float(), self.factor, self.taps.to(dev), zero, one, 4, clamp=False)
        return L.nhwc_to_nchw(y, 3)


class FaceParser(nn.Module):
    def __init__(self, seg_ckpt, size=1024, device="cuda"):
        super().__init__()
        self.seg_ckpt = seg_ckpt
        self.size = size
        self.device = device
        self.load_segmentation_network()
        self.load_downsampling()
        self._lut12 = None

    def load_downsampling(self):
        self.downsample = BicubicDownSample(factor=self.size // 512)
        self.downsample_256 = BicubicDownSample(factor=self.size // 256)

    def load_segmentation_network(self):
        self.seg = BiSeNet(n_classes=19)
        self.seg.to(self.device)
        if self.seg_ckpt is not None:
            self.seg.load_state_dict(torch.load(self.seg_ckpt, map_location=self.device))
        for param in self.seg.parameters():
            param.requires_grad = False
        self.seg.eval()

    def _consts(self, dev):
        """Filter taps / normalisation constants on `dev`, uploaded once per device (three tiny H2D copies per call otherwise, which also
        cannot be captured into a CUDA graph)."""
        cache = self.__dict__.setdefault("_const_cache", {})
        key = str(dev)
        if key not in cache:
            cache[key] = (self.downsample.taps.to(dev), seg_mean.reshape(3).to(dev), seg_std.reshape(3).to(dev))
        return cache[key]

    def preprocess_tensor(self, im01: torch.Tensor) -> torch.Tensor:
        """[B,3,S,S] in [0,1] on the GPU, S >= 512 -> normalised NHWC [B,S/f,S/f,RGB_PAD] with the FIXED factor f = self.size // 512
        of face_parsing_demo.py:138-139,151-156 (a 1024^2 parser given a 512^2 image parses at 256^2, like the reference)."""
        taps, mean, std = self._consts(im01.device)
        s, f = im01.shape[-1], self.downsample.factor
        if im01.shape[-2] != s or s < 512:
            raise L.E4SError("inputs smaller than 512 (or not square) go through preprocess_img (PIL resize)")
        if s % f or (s // f) % 32:
            raise L.E4SError(f"a {s}x{s} input down-sampled by {f} must be a multiple of 32 for BiSeNet")
        return L.bicubic_down_norm(im01.contiguous().float(), f, taps, mean, std, RGB_PAD)

    def preprocess_img(self, img):
        """PIL image -> normalised NHWC tensor."""
        import torchvision
        from PIL import Image
        if img.size[0] >= 512:
            im = torchvision.transforms.ToTensor()(img)[:3].unsqueeze(0).to(self.device)
            return self.preprocess_tensor(im)
        im = img.resize((512, 512), Image.BILINEAR)
        im = torchvision.transforms.ToTensor()(im)[:3].unsqueeze(0).to(self.device)
        im = (im.clamp(0, 1) - seg_mean.to(im.device)) / seg_std.to(im.device)
        return L.nchw_to_nhwc(im.contiguous(), RGB_PAD)

    @torch.no_grad()
    def parse_batch(self, im01: torch.Tensor, convert_to_seg12: bool = True) -> torch.Tensor:
        """Batched GPU entry point: [B,3,S,S] in [0,1] -> u8 labels [B,S/f,S/f] (512^2 for S = self.size; 12-region ids by default)."""
        x = self.preprocess_tensor(im01)
        hw = (x.shape[1], x.shape[2])
        lut = None
        if convert_to_seg12:
            if self._lut12 is None or self._lut12.device != x.device:
                self._lut12 = torch.from_numpy(SEG19_TO_SEG12.copy()).to(x.device)
            lut = self._lut12
        return self.seg.labels(x, hw, lut)

    @torch.no_grad()
    def forward(self, img):
        """PIL image -> [512,512] int64 label map of 19-class ids on the device (face_parsing_demo.py:162-176)."""
        x = self.preprocess_img(img)
        return self.seg.labels(x, (x.shape[1], x.shape[2]), None)[0].long()


def init_faceParsing_pretrained_model(ckpt_path):
    parser = FaceParser(seg_ckpt=ckpt_path)
    print("Load faceParsing pre-traiend model success!")
    return parser


def faceParsing_demo(model, img, convert_to_seg12=True):
    with torch.no_grad():
        seg = model(img).cpu().numpy().astype(np.uint8)
    if convert_to_seg12:
        seg = ffhq_masks_to_faceParser_mask_detailed(seg)
    return seg

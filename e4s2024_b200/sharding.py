"""Batch sharding of the swap hot path over the GPUs of one box.

Samples are independent (InstanceNorm is per sample, BatchNorm is in eval form, noise buffers are broadcast
constants), so the path shards by batch with NO data-path collective; the only exchange is one NCCL
all-gather of the output images (and label maps) over NVLink / NVSwitch.  One process per GPU, launched
with torchrun; the same code runs with the `gloo` backend on CPU for the host-logic tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_batch(local: torch.Tensor, total: Optional[int] = None, group=None) -> torch.Tensor:
    """Concatenate per-rank batches along dim 0 on every rank.  Equal shards take the single in-place
    all_gather_into_tensor (one NCCL ring/NVLS collective); ragged shards are padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    if total is not None and total % world == 0:
        counts = [total // world] * world
    else:
        dist.all_gather(sizes, n_local, group=group)
        counts = [int(s.item()) for s in sizes]
    if len(set(counts)) == 1:
        out = torch.empty((counts[0] * world,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    mx = max(counts)
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    buf = torch.empty((mx * world,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


class AsyncGather:
    """The path's only exchange, off the critical path: all-gather of this rank's outputs into a ring of `depth` pre-allocated
    buffers, asynchronously, so that the gather of step i rides under the kernels of step i+1 (NVLink / NVSwitch traffic does not
    depend on the next step's inputs).  The outputs leave as what the pipelines consume after the generator -- uint8 HWC images
    (tensor2im) and u8 label maps: 54 MB per rank and step at 16 faces instead of the 201 MB of the fp32 NCHW images, so the NCCL
    kernels take 4x less HBM bandwidth and time from the persistent convolution kernels they overlap."""

    def __init__(self, world: int, shapes_dtypes, device, depth: int = 2, group=None):
        self.world, self.group, self.depth = world, group, depth
        self.bufs = [[torch.empty((world * s[0],) + tuple(s[1:]), device=device, dtype=dt) for s, dt in shapes_dtypes] for _ in range(depth)]
        self.inflight = []
        self.n = 0

    def bytes_per_rank(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.bufs[0]) // self.world

    def submit(self, tensors):
        """tensors: this rank's outputs (same shapes as given to __init__) -> the gathered buffers (complete after wait())."""
        if len(self.inflight) == self.depth:             # the buffer set about to be reused: its gathers must have completed
            for w in self.inflight.pop(0):
                w.wait()
        bufs = self.bufs[self.n % self.depth]
        self.n += 1
        self.inflight.append([dist.all_gather_into_tensor(b, t.contiguous(), group=self.group, async_op=True) for b, t in zip(bufs, tensors)])
        return bufs

    def wait(self):
        while self.inflight:
            for w in self.inflight.pop(0):
                w.wait()


class SwapHotPath:
    """parse -> one-hot -> encode -> regional styles -> synthesise for this rank's shard of a batch
    (BASELINE.json config 5).  `net` is a Net3, `parser` a FaceParser; both hold replicated weights."""

    def __init__(self, net, parser, num_seg_cls: int = 12):
        self.net, self.parser, self.k = net, parser, num_seg_cls

    @torch.no_grad()
    def run_shard(self, img: torch.Tensor, randomize_noise: bool = False, img01: Optional[torch.Tensor] = None):
        """img [b,3,1024,1024] in [-1,1] on this rank's GPU -> (images [b,3,S,S], labels u8 [b,512,512]).
        img01: the same images in [0,1] (what the parser reads) when the caller already has them."""
        from . import _lib as L
        labels = self.parser.parse_batch((img + 1) * 0.5 if img01 is None else img01)
        mask = L.labels_to_onehot(labels, self.k)
        out, _ = self.net(img, mask, randomize_noise=randomize_noise)
        return out, labels

    @torch.no_grad()
    def run_shard_u8(self, img_u8: torch.Tensor, randomize_noise: bool = False):
        """The hand-off formats of the pipelines on both sides: uint8 HWC images [b,1024,1024,3] in (what PIL / cv2 hold; TO_TENSOR and
        NORMALIZE run on the device with the reference's arithmetic) -> (uint8 HWC images [b,S,S,3] = tensor2im of the synthesis,
        labels u8 [b,512,512])."""
        from . import _lib as L
        img01, img = L.im2tensor(img_u8)                 # ToTensor (parser input) and ToTensor + Normalize(0.5, 0.5) (Net3 input)
        out, labels = self.run_shard(img, randomize_noise, img01=img01)
        return L.tensor2im_u8(out, True), labels

    @torch.no_grad()
    def __call__(self, img_global_or_local: torch.Tensor, sharded_input: bool = True, randomize_noise: bool = False):
        """Every rank returns the full-batch images and label maps (all-gathered).  uint8 HWC input selects the u8 hand-off form
        (u8 images out: a quarter of the bytes in the all-gather), float NCHW input the reference modules' tensor form."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        x = img_global_or_local
        total = None
        if not sharded_input:
            total = x.shape[0]
            lo, hi = shard_range(total, rank, world)
            x = x[lo:hi]
        out, labels = self.run_shard_u8(x, randomize_noise) if x.dtype == torch.uint8 else self.run_shard(x, randomize_noise)
        return all_gather_batch(out, total), all_gather_batch(labels, total)

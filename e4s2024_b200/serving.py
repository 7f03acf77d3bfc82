"""Host <-> device streaming around the hot path: overlap the PCIe copies of batch i+1 / i-1 with the kernels of batch i.

The reference pipelines (face_swap_video_pipeline.py:430-443) run batch 1 and copy synchronously; at 16 faces per
batch the one-hot mask (201 MB) and the images (201 MB) cost as much PCIe time as the whole synthesis costs kernel time,
so a serving loop has to double-buffer them.  This module is plumbing only (pinned buffers, three CUDA streams,
events); every computation stays inside the callable it is given (the drop-in modules -> libe4s_b200.so).

    pipe = HostPipeline(lambda lat, msk: G([lat], None, msk, input_is_latent=True, randomize_noise=False)[0], device)
    for lat_h, msk_h, out_h in batches:        # pinned host tensors
        pipe.submit((lat_h, msk_h), out_h)
    pipe.drain()
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class HostPipeline:
    """Double-buffered H2D -> fn -> D2H.  `submit` returns immediately; results are complete after `drain()` (or after
    `depth` further submits).  Inputs must be pinned host tensors for the copies to be asynchronous."""

    def __init__(self, fn: Callable[..., torch.Tensor], device: torch.device, depth: int = 2):
        if device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device (there is no CPU path)")
        self.fn, self.dev, self.depth = fn, device, depth
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self.slots = [None] * depth            # device input buffers per slot
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # inputs of the slot are resident
        self.ev_used = [torch.cuda.Event() for _ in range(depth)]     # fn has consumed the slot's inputs
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # result of the slot is on the host
        self.pending: list = [None] * depth    # (inputs_host, out_host) staged for the slot, not yet computed
        self.n_staged = 0
        self.n_run = 0

    def _stage(self, inputs: Sequence[torch.Tensor]):
        """H2D of one batch into the next slot on the copy-in stream."""
        slot = self.n_staged % self.depth
        if self.slots[slot] is None:
            self.slots[slot] = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in inputs]
        with torch.cuda.stream(self.s_in):
            if self.n_staged >= self.depth:
                self.s_in.wait_event(self.ev_used[slot])       # the previous user of this slot has read its inputs
            for d, h in zip(self.slots[slot], inputs):
                d.copy_(h, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        self.n_staged += 1
        return slot

    def submit(self, inputs: Sequence[torch.Tensor], out_host: Optional[torch.Tensor]):
        """Queue one batch.  The batch submitted one call earlier is computed now (its copy overlapped that call's kernels)."""
        slot = self._stage(inputs)
        self.pending[slot] = out_host
        if self.n_staged - self.n_run >= self.depth:
            self._run_one()

    def _run_one(self):
        slot = self.n_run % self.depth
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ev_in[slot])
        res = self.fn(*self.slots[slot])
        self.ev_used[slot].record(cur)
        out_host = self.pending[slot]
        if out_host is not None:
            # one result tensor -> one pinned host tensor, or a tuple of results -> a tuple of host tensors (images + label maps)
            pairs = list(zip(res, out_host)) if isinstance(out_host, (tuple, list)) else [(res, out_host)]
            self.s_out.wait_event(self.ev_used[slot])
            with torch.cuda.stream(self.s_out):
                for d, h in pairs:
                    d.record_stream(self.s_out)
                    h.copy_(d, non_blocking=True)
                self.ev_out[slot].record(self.s_out)
        self.n_run += 1

    def drain(self):
        while self.n_run < self.n_staged:
            self._run_one()
        self.s_out.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()


class GraphedGenerator:
    """One CUDA graph of `Generator.forward` for a fixed (batch, regions, mask size): ~40 kernel launches, the label / job-list
    kernels and the style-table GEMMs replay as ONE launch from the host.  Possible because a forward contains no host wait: the
    region-job decisions are device-side launch predicates and the one-hot check of the mask is read AFTER the replay.

        gg = GraphedGenerator(G, batch=16, regions=12, mask_hw=(512, 512))
        img = gg(latent, mask)        # img is the graph's static output buffer: consume (or copy) it before the next call

    Noise: the registered buffers (`randomize_noise=False`), as the swap pipelines run the generator.
    Inputs must have exactly the captured shapes (no broadcasting into the static buffers)."""

    def __init__(self, G, batch: int, regions: int, mask_hw=(512, 512), device: Optional[torch.device] = None, warmup: int = 2):
        dev = device or next(G.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedGenerator needs a CUDA device (there is no CPU path)")
        self.G = G
        self.latent = torch.zeros(batch, regions, G.n_latent, G.style_dim, device=dev)
        self.mask = torch.zeros(batch, regions, *mask_hw, device=dev)
        self.mask[:, 0] = 1.0                                        # a valid one-hot mask for the warm-up / capture runs
        self.flag = torch.zeros(16, dtype=torch.int32).pin_memory()  # [0] = "mask was not one-hot" count, written by every replay
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                # warm-up off the default stream: packs weights, sets attributes
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.image, self.feats = self._run()

    def _run(self):
        img, _, feats = self.G([self.latent], None, self.mask, input_is_latent=True, randomize_noise=False, _host_flag=self.flag)
        return img, feats

    def __call__(self, latent: torch.Tensor, mask: torch.Tensor, check: bool = True) -> torch.Tensor:
        """check=True (default): wait for the replay and verify the mask was one-hot; a soft / overlapping / empty mask is re-run
        through `Generator.forward` (generic per-region path), so the result is always the reference's.  check=False returns without
        waiting (pipelined serving): the caller MUST call `verify()` before trusting the image."""
        if tuple(mask.shape) != tuple(self.mask.shape):
            raise ValueError(f"GraphedGenerator was captured for masks of shape {tuple(self.mask.shape)}, got {tuple(mask.shape)}")
        if latent.dim() != 4 or latent.shape[:2] != self.latent.shape[:2] or latent.shape[3] != self.latent.shape[3] or \
                latent.shape[2] < self.latent.shape[2]:
            raise ValueError(f"GraphedGenerator was captured for latents of shape {tuple(self.latent.shape)}, got {tuple(latent.shape)}")
        self.latent.copy_(latent[:, :, :self.latent.shape[2]], non_blocking=True)    # like forward: extra W+ layers are ignored
        self.mask.copy_(mask, non_blocking=True)
        self.graph.replay()
        if check and not self.verify():
            img, _, _ = self.G([self.latent], None, self.mask, input_is_latent=True, randomize_noise=False)
            return img
        return self.image

    def verify(self) -> bool:
        """Waits for the last replay; False if its mask was not one-hot (the graph's image is then NOT the reference's result)."""
        torch.cuda.current_stream(self.latent.device).synchronize()
        return int(self.flag[0]) == 0



class GraphedSwapPath:
    """One CUDA graph of the whole hot path for a fixed batch: uint8 HWC images in -> (uint8 HWC images, u8 label maps) out
    (`sharding.SwapHotPath.run_shard_u8`: TO_TENSOR / NORMALIZE, bicubic + BiSeNet + argmax / LUT, one-hot, encoder, LocalMLPs, generator,
    tensor2im) -- ~250 kernel launches replayed as ONE launch from the host.  The reference pipelines call the path frame by frame
    (batch 1), where the step is bound by the host's launch rate, not by the GPU.  Possible because the path contains no host wait: every
    data-dependent choice (region jobs, rows per cell) is a device-side launch predicate, and the mask comes from the parser's own label
    map, so it is one-hot by construction.

        gs = GraphedSwapPath(hot, batch=1)
        img_u8, labels = gs(frame_u8)        # the graph's static output buffers: consume (or copy) them before the next call
    """

    def __init__(self, hot, batch: int = 1, size: int = 1024, device: Optional[torch.device] = None, warmup: int = 2):
        dev = device or next(hot.net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedSwapPath needs a CUDA device (there is no CPU path)")
        if hot.net.training:
            raise RuntimeError("GraphedSwapPath captures the inference path: call net.eval() first")
        self.hot = hot
        self.inp = torch.zeros(batch, size, size, 3, dtype=torch.uint8, device=dev)
        self.flag = torch.zeros(16, dtype=torch.int32).pin_memory()
        G = hot.net.G
        G._capture_host_flag = self.flag          # no event wait / host read inside the forward (Generator._forward)
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                            # warm-up off the default stream: packs weights, sets attributes, sizes workspaces
                for _ in range(warmup):
                    hot.run_shard_u8(self.inp)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.image, self.labels = hot.run_shard_u8(self.inp)
        finally:
            G._capture_host_flag = None

    def __call__(self, img_u8: torch.Tensor):
        if tuple(img_u8.shape) != tuple(self.inp.shape) or img_u8.dtype != torch.uint8:
            raise ValueError(f"GraphedSwapPath was captured for uint8 images of shape {tuple(self.inp.shape)}, got {img_u8.dtype} {tuple(img_u8.shape)}")
        self.inp.copy_(img_u8, non_blocking=True)
        self.graph.replay()
        return self.image, self.labels

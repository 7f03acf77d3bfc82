"""CPU, world_size 2 over gloo: the batch-sharding bookkeeping and the single all-gather exchange.
The compute inside each shard is stubbed with a deterministic per-sample function so the test isolates the
N>1 host path: sharded == unsharded, equal and ragged shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from e4s2024_b200.sharding import all_gather_batch, shard_range


def test_shard_range_partition():
    for n in (0, 1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 2, 3)
        f = lambda x: x * 2 + 1                                   # per-sample "hot path"
        lo, hi = shard_range(total, rank, world)
        out = all_gather_batch(f(full[lo:hi]), total)
        ok = torch.equal(out, f(full))
        labels = all_gather_batch(full[lo:hi, 0, :1].to(torch.uint8))   # ragged-safe path without `total`
        ok = ok and labels.shape[0] == total
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_equals_unsharded_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


class _StubParser:
    def parse_batch(self, img01):
        return (img01.mean(dim=1)[:, ::2, ::2] * 11.999).to(torch.uint8).contiguous()     # [b, h/2, w/2] labels in 0..11


class _StubNet:
    def __call__(self, img, mask, randomize_noise=False):
        return img * mask.sum(dim=1, keepdim=True).repeat_interleave(2, 2).repeat_interleave(2, 3) + 1.0, None


def _worker_path(rank, world, port, total, q):
    """SwapHotPath.__call__ (shard -> run -> all-gather of images and label maps) and AsyncGather with stub networks: the N > 1 host
    path of the bench, on gloo."""
    from e4s2024_b200.sharding import AsyncGather, SwapHotPath
    from tests import cpu_emul
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        full = torch.rand(total, 3, 8, 8, generator=g) * 2 - 1
        hot = SwapHotPath(_StubNet(), _StubParser(), 12)
        with cpu_emul.emulated():
            images, labels = hot(full, sharded_input=False)
            ref_img, ref_lab = hot.run_shard(full)
        ok = torch.equal(images, ref_img) and torch.equal(labels, ref_lab)
        # asynchronous, double-buffered gather of (uint8 images, label maps): three steps through two buffer sets
        ag = AsyncGather(world, [((2, 4, 4, 3), torch.uint8), ((2, 5), torch.uint8)], torch.device("cpu"))
        outs = []
        for step in range(3):
            a = torch.full((2, 4, 4, 3), 10 * step + rank, dtype=torch.uint8)
            b = torch.full((2, 5), 100 + 10 * step + rank, dtype=torch.uint8)
            bufs = ag.submit([a, b])
            if step == 2:
                outs = bufs
        ag.wait()
        for r in range(world):
            ok = ok and bool((outs[0][2 * r: 2 * r + 2] == 20 + r).all()) and bool((outs[1][2 * r: 2 * r + 2] == 120 + r).all())
        ok = ok and ag.bytes_per_rank() == 2 * 4 * 4 * 3 + 2 * 5
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_swap_hot_path_and_async_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_path, args=(r, 2, port, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

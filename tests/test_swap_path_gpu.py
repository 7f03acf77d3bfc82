"""GPU: the full swap hot path through its public entry (sharding.SwapHotPath, BASELINE.json configs[4] per-GPU shard) against the
oracle chain, the uint8 hand-off form, and the masks a real face produces (curved boundaries, per-pixel noise) at 1024^2.
The 2-rank NCCL test (sharded == unsharded, bit for bit) needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_swap_path_gpu.py`."""
import os
import socket

import numpy as np
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc

pytestmark = pytest.mark.gpu
K = 12


def _path(out_size=1024, dev="cuda"):
    from oracle.ref_shims import net3_opts
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    from e4s2024_b200.networks import Net3
    from e4s2024_b200.sharding import SwapHotPath
    net = Net3(net3_opts(out_size=out_size, remaining_layer_idx=13))
    sd_net = synth.synth_module_weights(net, seed=9)
    net = net.to(dev)
    la = synth.randn("net3.latent_avg", (18, 512), 9, 0.1)
    net.latent_avg = la.to(dev)
    parser = FaceParser(seg_ckpt=None, size=1024, device=dev)
    sd_seg = synth.synth_module_weights(parser.seg, seed=10)
    parser.seg.to(dev)
    sds = {"net": {k: v.cpu() for k, v in sd_net.items()}, "seg": {k: v.cpu() for k, v in sd_seg.items()}, "latent_avg": la}
    return SwapHotPath(net, parser, K), sds


def test_swap_hot_path_vs_oracle_chain():
    """run_shard_u8 (uint8 in -> parse -> one-hot -> Net3 -> uint8 out) for 2 faces vs the oracle executing the reference's chain on
    the CPU: label maps equal up to the tie rule, images within 1e-3 (fp32) / 1 grey level (uint8), sample alone == sample in batch."""
    hot, sds = _path(1024)
    img_u8 = synth.smooth_image_u8("swappath.img", 2, 1024, 31)
    out_u8, labels = hot.run_shard_u8(img_u8.cuda())
    assert out_u8.shape == (2, 1024, 1024, 3) and out_u8.dtype == torch.uint8 and labels.shape == (2, 512, 512)
    x01, x = orc.to_tensor_normalize(img_u8[:1])
    logits = orc.bisenet_forward(sds["seg"], orc.parser_preprocess(x01, 1024))[0]
    ref_lab = torch.from_numpy(orc.SEG19_TO_SEG12)[logits.argmax(1).long()]
    bad = ref_lab != labels[:1].cpu()
    top2 = torch.topk(logits, 2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    scale = float(logits.abs().max())
    worst = float(margin[bad].max()) if bad.any() else 0.0
    print(f"swap path labels: {int(bad.sum())} of {bad.numel()} differ, largest oracle margin at a flip {worst:.3e} (tie threshold {2e-5 * scale:.3e})")
    assert int(bad.sum()) <= 8 and worst < 2e-5 * scale
    # image parity on the SAME mask (the label map was compared above)
    mask = orc.label_to_onehot(labels[:1].cpu().long()[:, None], K)
    ref_img = orc.net3_forward(sds["net"], x, mask, sds["latent_avg"], out_size=1024, remaining_layer_idx=13)[0]
    x01_d, x_d = [t.cuda() for t in (x01, x)]
    img_f, lab_f = hot.run_shard(x_d, img01=x01_d)
    d = float((img_f.cpu() - ref_img).abs().max())
    print(f"swap path image: max|diff| {d:.3e} (range {float(ref_img.abs().max()):.2f})")
    assert d < 1e-3
    assert torch.equal(lab_f, labels[:1])                                   # batch invariance of the parser
    ref_u8 = orc.tensor2im_u8(ref_img)
    du8 = int((out_u8[:1].cpu().int() - ref_u8.int()).abs().max())
    assert du8 <= 1, du8
    solo_u8, _ = hot.run_shard_u8(img_u8[1:].cuda())
    assert torch.equal(solo_u8[0], out_u8[1])                                # => sharded == unsharded


@pytest.mark.parametrize("kind", ["face", "noise"])
def test_generator_1024_hard_masks_vs_oracle(kind):
    """Curved region boundaries (multi-region tiles on every masked layer: the wide kernel's per-phase fallback and multi-job tiles of
    the halo kernel inside a full forward) and per-pixel noise labels (all 12 regions in every tile) at 1024^2, <= 1e-3 vs the oracle."""
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(1024, 512, 8, split_layer_idx=5, remaining_layer_idx=13)
    sd = synth.synth_module_weights(G, seed=2)
    G = G.cuda().eval().requires_grad_(False)
    latent = synth.randn(f"hard.{kind}.latent", (2, K, 18, 512), 41)
    mask = synth.onehot(synth.make_labels(kind, 2, K, 512, seed=41), K)
    img, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    ref, _ = orc.generator_forward({k: v.cpu() for k, v in sd.items()}, 1024, latent[:1], mask[:1], split_layer_idx=5, remaining_layer_idx=13)
    d = float((img[:1].cpu() - ref).abs().max())
    print(f"1024^2 {kind} masks: image max|diff| vs oracle {d:.3e} (range {float(ref.abs().max()):.2f})")
    assert d < 1e-3
    solo, _, _ = G([latent[1:].cuda()], None, mask[1:].cuda(), input_is_latent=True, randomize_noise=False)
    assert torch.equal(solo[0], img[1])


def test_net3_256_face_masks_vs_oracle():
    """Encoder + MLPs + generator with curved masks (masked-mean pooling over curved regions, empty regions -> zero style)."""
    hot, sds = _path(256)
    img = synth.smooth_image("net3face.img", 2, 1024, 43)
    mask = synth.onehot(synth.make_labels("face", 2, K, 512, seed=43), K)
    out, _ = hot.net(img.cuda(), mask.cuda(), randomize_noise=False)
    ref = orc.net3_forward(sds["net"], img[:1], mask[:1], sds["latent_avg"], out_size=256, remaining_layer_idx=13)[0]
    d = float((out[:1].cpu() - ref).abs().max())
    print(f"Net3 256^2 face masks: image max|diff| {d:.3e} (range {float(ref.abs().max()):.2f})")
    assert d < 1e-3


# ---- two ranks over NCCL: sharded == unsharded ---------------------------------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        hot, _ = _path(256, dev=f"cuda:{rank}")
        full = synth.smooth_image_u8("nccl.img", 4, 1024, 51).to(f"cuda:{rank}")
        images, labels = hot(full, sharded_input=False)              # every rank: all-gathered uint8 images + label maps
        ok = images.shape == (4, 256, 256, 3) and labels.shape == (4, 512, 512)
        if rank == 0:
            ref_img, ref_lab = hot.run_shard_u8(full)                # the whole batch on one GPU
            ok = ok and torch.equal(images, ref_img) and torch.equal(labels, ref_lab)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_equals_unsharded_nccl_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(0, True), (1, True)]

"""GPU: the double-buffered host pipeline (e4s2024_b200/serving.py) returns, batch for batch, exactly what a direct
Generator.forward on device tensors returns -- overlapping the PCIe copies must not change a bit."""
import pytest
import torch

from e4s2024_b200 import synth

pytestmark = pytest.mark.gpu


def test_host_pipeline_bit_exact():
    from e4s2024_b200.serving import HostPipeline
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(256, 512, 8, split_layer_idx=5, remaining_layer_idx=13)
    synth.synth_module_weights(G, seed=6)
    G = G.cuda().eval()
    dev = torch.device("cuda", torch.cuda.current_device())

    def fn(lat, msk):
        return G([lat], None, msk, input_is_latent=True, randomize_noise=False)[0]

    batches = []
    for i in range(5):
        lat = synth.randn(f"pipe.latent{i}", (2, 12, 18, 512), 30 + i).pin_memory()
        msk = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=30 + i), 12).pin_memory()
        batches.append((lat, msk, torch.empty(2, 3, 256, 256).pin_memory()))
    pipe = HostPipeline(fn, dev)
    for lat, msk, out in batches:
        pipe.submit((lat, msk), out)
    pipe.drain()
    for lat, msk, out in batches:
        ref = fn(lat.cuda(), msk.cuda()).cpu()
        assert torch.equal(out, ref)


def test_graphed_generator_bit_exact():
    """One CUDA graph of the whole forward replays to exactly the eager result, for several inputs."""
    from e4s2024_b200.serving import GraphedGenerator
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(256, 512, 8, split_layer_idx=5, remaining_layer_idx=9)
    synth.synth_module_weights(G, seed=6)
    G = G.cuda().eval()
    gg = GraphedGenerator(G, batch=2, regions=12, mask_hw=(512, 512))
    for i in range(3):
        lat = synth.randn(f"graph.latent{i}", (2, 12, 18, 512), 80 + i).cuda()
        msk = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=80 + i), 12).cuda()
        out = gg(lat, msk, check=True).clone()
        ref = G([lat], None, msk, input_is_latent=True, randomize_noise=False)[0]
        assert torch.equal(out, ref), float((out - ref).abs().max())
    soft = msk.clone()
    soft[:, :, :8, :8] = 0.25                                        # not one-hot: never a wrong image -- the default call falls
    out = gg(lat, soft)                                              # back to Generator.forward's generic per-region path
    ref = G([lat], None, soft, input_is_latent=True, randomize_noise=False)[0]
    assert torch.equal(out, ref)
    gg(lat, soft, check=False)                                       # pipelined form: the caller must ask
    assert gg.verify() is False
    gg(lat, msk, check=False)
    assert gg.verify() is True
    with pytest.raises(ValueError):                                  # no silent broadcast of a batch-1 input into the static buffers
        gg(lat[:1], msk[:1])


def test_graphed_swap_path_bit_exact():
    """serving.GraphedSwapPath: the whole hot path (parse -> one-hot -> Net3 -> tensor2im) replayed as one CUDA graph gives the bytes of
    the eager path, for the capture input and for new inputs (the region job / row lists are rebuilt on the device at every replay)."""
    from e4s2024_b200 import synth
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    from e4s2024_b200.networks import Net3
    from e4s2024_b200.serving import GraphedSwapPath
    from e4s2024_b200.sharding import SwapHotPath
    from oracle.ref_shims import net3_opts
    net = Net3(net3_opts(out_size=256, remaining_layer_idx=13))
    synth.synth_module_weights(net, seed=9)
    net = net.cuda().eval()
    net.latent_avg = synth.randn("gs.latent_avg", (18, 512), 9, 0.1).cuda()
    parser = FaceParser(seg_ckpt=None, size=1024, device="cuda")
    synth.synth_module_weights(parser.seg, seed=10)
    parser.seg.cuda()
    hot = SwapHotPath(net, parser, 12)
    gs = GraphedSwapPath(hot, batch=2)
    for seed in (3, 4):
        u8 = synth.smooth_image_u8("gs.img", 2, 1024, seed).cuda()
        img_g, lab_g = gs(u8)
        img_g, lab_g = img_g.clone(), lab_g.clone()
        img_e, lab_e = hot.run_shard_u8(u8)
        torch.cuda.synchronize()
        assert torch.equal(lab_g, lab_e) and torch.equal(img_g, img_e)
    assert int(gs.flag[0]) == 0

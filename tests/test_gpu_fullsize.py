"""GPU: parity at BASELINE.json's full sizes (per-pixel <= 1e-3 vs the oracle at 1024^2) and the
size-independent properties the domain offers: batch invariance (sample i of a batch == the same sample
run alone, bit for bit -> sharded == unsharded), determinism, mask-region locality."""
import numpy as np
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc

pytestmark = pytest.mark.gpu


def _gen(size, rl, split, seed=2):
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(size, 512, 8, split_layer_idx=split, remaining_layer_idx=rl)
    sd = synth.synth_module_weights(G, seed=seed)
    return G.cuda().eval(), {k: v.cpu() for k, v in sd.items()}


def test_generator_256_vs_oracle_both_engines():
    from e4s2024_b200 import engine as E
    G, sd = _gen(256, 13, 5)
    latent = synth.randn("t256.latent", (2, 12, 18, 512), 3)
    mask = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=3), 12)
    ref, rinter = orc.generator_forward(sd, 256, latent, mask, split_layer_idx=5, remaining_layer_idx=13)
    for eng in ("f32", "tc") if E.tc_available() else ("f32",):
        E.set_conv_engine(eng)
        img, _, inter = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
        d = float((img.cpu() - ref).abs().max())
        di = float((inter.cpu() - rinter).abs().max())
        print(f"engine {eng}: image max|diff| {d:.3e} (range {float(ref.abs().max()):.2f}), inter {di:.3e}")
        assert d < 1e-3 and di < 1e-3, (eng, d, di)
    E.set_conv_engine("tc")


def test_generator_1024_vs_oracle_and_batch_invariance():
    G, sd = _gen(1024, 13, 5)
    latent = synth.randn("t1024.latent", (2, 12, 18, 512), 4)
    mask = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=4), 12)
    img, _, inter = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    assert img.shape == (2, 3, 1024, 1024) and inter.shape == (2, 512, 16, 16)
    ref, _ = orc.generator_forward(sd, 1024, latent[:1], mask[:1], split_layer_idx=5, remaining_layer_idx=13)
    d = float((img[:1].cpu() - ref).abs().max())
    print(f"1024^2 image max|diff| vs oracle {d:.3e} (range {float(ref.abs().max()):.2f})")
    assert d < 1e-3
    # batch invariance: sample 1 alone == sample 1 inside the batch, bit for bit (=> shard-equivalence)
    solo, _, _ = G([latent[1:].cuda()], None, mask[1:].cuda(), input_is_latent=True, randomize_noise=False)
    assert torch.equal(solo[0], img[1])
    again, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    assert torch.equal(again, img)                                   # deterministic


def test_region_locality_and_soft_mask_path():
    """Changing the style of one region only changes... everything downstream of its pixels; but a region with
    no pixels must have no influence at all.  And a one-hot mask pushed through the generic float-mask path
    (mask * 1.0 with a tiny perturbation removed) must agree with the label fast path."""
    from e4s2024_b200 import engine as E
    G, _ = _gen(64, 18, 7)
    latent = synth.randn("loc.latent", (1, 4, 10, 512), 5)
    lab = synth.blocky_labels(1, 3, 64, cells=8, seed=5)            # region 3 never appears
    mask = synth.onehot(lab, 4)
    a, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    lat2 = latent.clone()
    lat2[:, 3] += 1.0
    b, _, _ = G([lat2.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    assert torch.equal(a, b)
    soft = mask * 0.5                                               # not one-hot -> per-region accumulate path
    c, _, _ = G([latent.cuda()], None, soft.cuda(), input_is_latent=True, randomize_noise=False)
    ctx = E.RegionCtx(soft.cuda())
    assert not ctx.onehot
    sd = {k: v.cpu() for k, v in G.state_dict().items()}
    ref, _ = orc.generator_forward(sd, 64, latent, soft, split_layer_idx=7, remaining_layer_idx=18)
    assert float((c.cpu() - ref).abs().max()) < 1e-3


def test_randomize_noise_and_explicit_noise():
    G, sd = _gen(32, 18, 7)
    latent = synth.randn("nz.latent", (2, 3, 8, 512), 6)
    mask = synth.onehot(synth.blocky_labels(2, 3, 32, cells=4, seed=6), 3)
    a, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=True)
    b, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=True)
    assert not torch.equal(a, b)                                   # fresh noise per call, like the reference
    chans = {4: 512, 8: 512, 16: 512, 32: 512}
    noise = [synth.randn("nz.0", (1, 512, 4, 4), 6)]
    for r in (8, 16, 32):
        noise += [synth.randn(f"nz.{r}a", (1, chans[r], r, r), 6), synth.randn(f"nz.{r}b", (1, chans[r], r, r), 6)]
    img, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, noise=[n.cuda() for n in noise])
    ref, _ = orc.generator_forward(sd, 32, latent, mask, noise)     # per-channel noise [1,C,r,r] (Face_swap_frontal.py:35-38)
    assert float((img.cpu() - ref).abs().max()) < 1e-3


def test_parser_batch_is_batch_invariant():
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    parser = FaceParser(seg_ckpt=None, size=1024, device="cuda")
    synth.synth_module_weights(parser.seg, seed=10)
    parser.seg.cuda()
    img01 = ((synth.smooth_image("pb.img", 3, 1024, 12) + 1) / 2).cuda()
    lab = parser.parse_batch(img01)
    assert lab.shape == (3, 512, 512) and lab.dtype == torch.uint8 and int(lab.max()) < 12
    assert torch.equal(parser.parse_batch(img01[1:2])[0], lab[1])


def test_net3_1024_vs_oracle():
    """BASELINE config 3 shape (encoder -> regional styles -> synthesis) at 1024^2, B=2 on the GPU vs the oracle."""
    from oracle.ref_shims import net3_opts
    from e4s2024_b200.networks import Net3
    net = Net3(net3_opts(out_size=1024, remaining_layer_idx=13))
    sd = synth.synth_module_weights(net, seed=9)
    net = net.cuda()
    la = synth.randn("net3.latent_avg", (18, 512), 9, 0.1)
    net.latent_avg = la.cuda()
    img = synth.smooth_image("net3full.img", 2, 1024, 13)
    mask = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=13), 12)
    out, inter = net(img.cuda(), mask.cuda(), randomize_noise=False)
    assert out.shape == (2, 3, 1024, 1024) and inter.shape == (2, 512, 16, 16)
    sdc = {k: v.cpu() for k, v in sd.items()}
    ro, _, rcodes, rvec = orc.net3_forward(sdc, img[:1], mask[:1], la, out_size=1024, remaining_layer_idx=13)
    vec, _ = net.get_style_vectors(img.cuda(), mask.cuda())
    dv = float((vec[:1].cpu() - rvec).abs().max())
    di = float((out[:1].cpu() - ro).abs().max())
    print(f"Net3 1024^2: style vectors max|diff| {dv:.3e} (scale {float(rvec.abs().max()):.2f}), image {di:.3e} (range {float(ro.abs().max()):.2f})")
    assert dv < 1e-3 and di < 1e-3
    solo, _ = net(img[1:].cuda(), mask[1:].cuda(), randomize_noise=False)
    assert torch.equal(solo[0], out[1])


def _bisenet_vs_oracle(engine, batch=4):
    from e4s2024_b200.face_parsing import resnet as rn
    from e4s2024_b200.face_parsing.model import BiSeNet
    seg = BiSeNet(19)
    sd = synth.synth_module_weights(seg, seed=10)
    seg = seg.cuda().eval()
    x = synth.randn("bise512.x", (batch, 3, 512, 512), 14)
    prev = rn.bisenet_engine()
    rn.set_bisenet_engine(engine)
    try:
        o, o16, o32 = seg(x.cuda())
    finally:
        rn.set_bisenet_engine(prev)
    ref = orc.bisenet_forward({k: v.cpu() for k, v in sd.items()}, x)[0]
    scale = float(ref.abs().max())
    d = float((o.cpu() - ref).abs().max())
    lab, rlab = o.argmax(1).cpu(), ref.argmax(1)
    bad = lab != rlab
    top2 = torch.topk(ref, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    worst = float(margin[bad].max()) if bad.any() else 0.0
    print(f"BiSeNet 512^2 B={batch} engine {engine}: logits max|diff| {d:.3e} (scale {scale:.1f}); label mismatches {int(bad.sum())} of "
          f"{bad.numel()}, largest oracle margin at a mismatch {worst:.3e}")
    return d, scale, int(bad.sum()), worst


@pytest.mark.parametrize("engine", ["tc16", "f32"])
def test_bisenet_512_argmax_vs_oracle(engine):
    """BASELINE config 4 shape at B=4: argmax label maps vs the oracle; mismatches only where the oracle's own top-2 margin is
    within fp32 reassociation noise of a tie.  The SAME bar for the default tensor-core mode (tcgen05, fp16 hi/lo split) and for
    the exact-fp32 CUDA-core engine that cross-checks it."""
    d, scale, nbad, worst = _bisenet_vs_oracle(engine)
    assert d < 3e-5 * scale
    assert nbad <= 16 and worst < 2e-5 * scale


def test_bisenet_bf16_split_is_not_enough():
    """Why BiSeNet does not share the generator's bf16 hi/lo format: its logits are 2-3x further from the oracle and label flips
    appear at margins fp32 arithmetic resolves.  (Documents the choice; only a loose bound is asserted.)"""
    d, scale, nbad, worst = _bisenet_vs_oracle("tc")
    assert d < 3e-4 * scale and nbad < 200


def test_fused_torgb_matches_separate_torgb(monkeypatch):
    """The ToRGB tail fused into the conv epilogue (256^2 layers of a 256^2 generator) against the stand-alone
    torgb kernel reading the stored feature map: same modulated 1x1 conv, bias and FIR-upsampled skip."""
    G, _ = _gen(256, 9, 5)            # rl=9: the 128^2 and 256^2 layers are un-masked -> fused tail on both resolutions
    latent = synth.randn("fuse.latent", (2, 12, 18, 512), 21)
    mask = synth.onehot(synth.blocky_labels(2, 12, 512, cells=32, seed=21), 12)
    fused, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    monkeypatch.setenv("E4S_FUSE_RGB", "0")
    plain, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    d = float((fused - plain).abs().max())
    print(f"fused vs separate ToRGB: max|diff| {d:.3e} (range {float(plain.abs().max()):.2f})")
    assert 0.0 < d < 2e-5 * max(float(plain.abs().max()), 1.0)      # different summation order, same numbers


def test_generator_256_all_layers_masked_and_small_mask():
    """remaining_layer_idx=18 (every StyledConv / ToRGB regional, the reference default), K=5 regions, a 128^2 mask:
    the masked 128^2 / 256^2 layers (cout 128 / 64) take the region-job halo kernel or the gather kernel."""
    G, sd = _gen(256, 18, 7, seed=5)
    latent = synth.randn("t256m.latent", (2, 5, 18, 512), 7)
    mask = synth.onehot(synth.blocky_labels(2, 5, 128, cells=8, seed=7), 5)
    ref, _ = orc.generator_forward(sd, 256, latent[:1], mask[:1], split_layer_idx=7, remaining_layer_idx=18)
    img, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    d = float((img[:1].cpu() - ref).abs().max())
    print(f"256^2 rl=18 K=5 mask 128^2: image max|diff| {d:.3e} (range {float(ref.abs().max()):.2f})")
    assert d < 1e-3
    solo, _, _ = G([latent[1:].cuda()], None, mask[1:].cuda(), input_is_latent=True, randomize_noise=False)
    assert torch.equal(solo[0], img[1])


@pytest.mark.parametrize("size,rl", [(128, 9), (512, 13), (512, 11)])
def test_generator_other_sizes_vs_oracle(size, rl):
    """Output sizes other than the two benchmarked ones: different sets of layers end up masked / un-masked (and therefore on
    different kernels: e.g. 512^2 with rl=11 runs 256^2 un-masked with cout=128 -> fused ToRGB at three resolutions)."""
    G, sd = _gen(size, rl, 5, seed=8)
    latent = synth.randn(f"t{size}.{rl}.latent", (1, 12, 18, 512), 11)
    mask = synth.onehot(synth.blocky_labels(1, 12, 512, cells=32, seed=11), 12)
    ref, _ = orc.generator_forward(sd, size, latent, mask, split_layer_idx=5, remaining_layer_idx=rl)
    img, _, _ = G([latent.cuda()], None, mask.cuda(), input_is_latent=True, randomize_noise=False)
    d = float((img.cpu() - ref).abs().max())
    print(f"{size}^2 rl={rl}: image max|diff| {d:.3e} (range {float(ref.abs().max()):.2f})")
    assert d < 1e-3

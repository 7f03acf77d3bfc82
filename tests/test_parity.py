"""Parity of the drop-in modules against the committed golden vectors (reference outputs) and the oracle.

Every case runs twice:
  * dev="emul": CPU, with the C-ABI wrappers replaced by tests/cpu_emul.py -> checks the HOST logic
    (packing, struct filling, offsets, layer order) without a GPU;
  * dev="cuda" (marked gpu): the real sm_100a kernels through libe4s_b200.so.
Tolerance: the north star's per-pixel 1e-3 on images; tighter (1e-4) on single ops.
"""
import contextlib

import numpy as np
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc
from tests import cpu_emul

T = torch.from_numpy
DEVS = ["emul", pytest.param("cuda", marks=pytest.mark.gpu)]


def ctx_for(dev):
    return cpu_emul.emulated() if dev == "emul" else contextlib.nullcontext()


def to(dev, *ts):
    d = "cpu" if dev == "emul" else "cuda"
    out = [t.to(d) if isinstance(t, torch.Tensor) else t for t in ts]
    return out[0] if len(out) == 1 else out


def load(mod, shapes_seed, dev):
    sd = mod.state_dict()
    new = synth.fill_state_dict({k: v.shape for k, v in sd.items()}, seed=shapes_seed)
    sd.update({k: v.to(sd[k].dtype) for k, v in new.items()})
    mod.load_state_dict(sd)
    return mod.to("cpu" if dev == "emul" else "cuda").eval()        # the pipelines run the modules in eval() mode


def maxdiff(a, b):
    return float((torch.as_tensor(a).detach().double().cpu() - torch.as_tensor(b).detach().double().cpu()).abs().max())


@pytest.mark.parametrize("dev", DEVS)
def test_upfirdn2d_and_bias_act(golden, dev):
    from e4s2024_b200 import _lib as L
    g = golden("upfirdn2d")
    with ctx_for(dev):
        x = to(dev, T(g["x"]))
        for name in ("blur", "up2", "down2", "crop", "up2down3"):
            up, down, p0, p1 = [int(v) for v in g[name + "_cfg"]]
            y = L.upfirdn2d(x, to(dev, T(g[name + "_kernel"])), up, down, p0, p1)
            assert maxdiff(y, g[name]) < 1e-5, name
        f = golden("fused_leaky_relu")
        y = L.bias_act(to(dev, T(f["x"])), to(dev, T(f["bias"])), 0.2, 2 ** 0.5)
        assert maxdiff(y, f["y"]) < 1e-6
        y = L.bias_act(to(dev, T(f["x"])), to(dev, T(f["bias"])), 0.1, 1.5)
        assert maxdiff(y, f["y_slope01"]) < 1e-6


@pytest.mark.gpu
def test_ops_public_api_cuda(golden):
    """The reference-facing op functions (same names/signature as models.stylegan2.op)."""
    from e4s2024_b200.stylegan2.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
    g = golden("upfirdn2d")
    y = upfirdn2d(T(g["x"]).cuda(), T(g["up2_kernel"]).cuda(), up=2, down=1, pad=(2, 1))
    assert maxdiff(y, g["up2"]) < 1e-5
    f = golden("fused_leaky_relu")
    assert maxdiff(fused_leaky_relu(T(f["x"]).cuda(), T(f["bias"]).cuda()), f["y"]) < 1e-6
    m = FusedLeakyReLU(6).cuda()
    m.bias.data.copy_(T(f["bias"]))
    assert maxdiff(m(T(f["x"]).cuda()).detach(), f["y"]) < 1e-6
    with pytest.raises(RuntimeError):
        fused_leaky_relu(T(f["x"]), T(f["bias"]))          # CPU tensors are rejected, like the reference op


@pytest.mark.parametrize("dev", DEVS)
@pytest.mark.parametrize("tag,k,demod,up", [("k3", 3, True, False), ("k3up", 3, True, True), ("k1nodemod", 1, False, False)])
def test_modulated_conv(golden, dev, tag, k, demod, up):
    from e4s2024_b200.stylegan2.model import ModulatedConv2d
    g = golden("modconv_" + tag)
    with ctx_for(dev):
        m = load(ModulatedConv2d(8, 16, k, 512, demodulate=demod, upsample=up), 3, dev)
        y = m(*to(dev, T(g["x"]), T(g["style"])))
    assert maxdiff(y, g["y"]) < 1e-4


@pytest.mark.parametrize("dev", DEVS)
@pytest.mark.parametrize("tag,up,seed", [("same", False, 4), ("up", True, 4), ("softmask", False, 6)])
def test_styled_conv(golden, dev, tag, up, seed):
    from e4s2024_b200.stylegan2.model import StyledConv
    g = golden("styledconv_" + tag)
    with ctx_for(dev):
        m = load(StyledConv(8, 16, 3, 512, upsample=up, mask_op=True), seed, dev)
        y = m(*to(dev, T(g["x"]), T(g["style"]), T(g["mask"])), noise=to(dev, T(g["noise"])))
    assert maxdiff(y, g["y"]) < 1e-4


@pytest.mark.parametrize("dev", DEVS)
def test_torgb(golden, dev):
    from e4s2024_b200.stylegan2.model import ToRGB
    g = golden("torgb")
    with ctx_for(dev):
        m = load(ToRGB(8, 512, upsample=True, mask_op=True), 5, dev)
        y = m(*to(dev, T(g["x"]), T(g["style"]), T(g["mask"]), T(g["skip"])))
        assert maxdiff(y, g["y"]) < 1e-4
        # soft mask -> generic accumulate path, checked against the oracle
        soft = torch.softmax(synth.randn("torgb.soft", (2, 5, 16, 16), 1), dim=1)
        ys = m(*to(dev, T(g["x"]), T(g["style"]), soft, T(g["skip"])))
        sd = {k: v.cpu() for k, v in m.state_dict().items()}
        yo = orc.to_rgb(T(g["x"]), T(g["style"]), soft, T(g["skip"]), sd, "", mask_op=True)
    assert maxdiff(ys, yo) < 1e-4


@pytest.mark.parametrize("dev", DEVS)
@pytest.mark.parametrize("tag", ["g32_rl18", "g64_rl5"])
def test_generator_small(golden, dev, tag):
    from e4s2024_b200.stylegan2.model import Generator
    g = golden("generator_" + tag)
    size, rl, split, K, seed = [int(v) for v in g["cfg"]]
    with ctx_for(dev):
        G = load(Generator(size, 512, 8, split_layer_idx=split, remaining_layer_idx=rl), seed, dev)
        mask = synth.onehot(T(g["labels"].astype(np.int64)), K)
        latent = synth.randn(f"{tag}.latent", (2, K, G.n_latent, 512), seed)
        img, lat, inter = G([to(dev, latent)], None, to(dev, mask), input_is_latent=True, randomize_noise=False)
    assert lat is None and img.shape == (2, 3, size, size)
    assert maxdiff(img, g["image"]) < 1e-3
    assert maxdiff(inter[:, ::16], g["inter"]) < 1e-3
    # eval() mode with parameters that require grad + autograd recording: the forward is the inference path, and backward fails with a
    # message that says why (not autograd's "does not require grad"); gradients exist in train() mode (tests/test_backward_gpu.py)
    from e4s2024_b200._lib import E4SError
    assert img.requires_grad
    with pytest.raises(E4SError, match="inference-only"):
        img.mean().backward()
    with torch.no_grad(), ctx_for(dev):
        img2 = G([to(dev, latent)], None, to(dev, mask), input_is_latent=True, randomize_noise=False)[0]
    assert not img2.requires_grad and maxdiff(img2, img) == 0.0


@pytest.mark.parametrize("dev", DEVS)
def test_encoder(golden, dev):
    from e4s2024_b200.encoders.psp_encoders import FSEncoder_PSP
    g = golden("encoder")
    with ctx_for(dev):
        enc = load(FSEncoder_PSP("ir_se", None), 8, dev)
        mask = synth.onehot(T(g["labels"].astype(np.int64)), 12)
        codes, struct = enc(*to(dev, T(g["x"]), mask))
    assert codes.shape == (2, 12, 1280) and struct.shape == (2, 512, 8, 8)
    assert float(struct.abs().max()) == 0.0 and float(codes[1, 3].abs().max()) == 0.0
    assert maxdiff(codes, g["codes"]) < 1e-3


@pytest.mark.parametrize("dev", DEVS)
def test_net3(golden, dev):
    from oracle.ref_shims import net3_opts
    from e4s2024_b200.networks import Net3
    g = golden("net3")
    out_size, rl, seed = [int(v) for v in g["cfg"]]
    with ctx_for(dev):
        net = load(Net3(net3_opts(out_size=out_size, remaining_layer_idx=rl)), seed, dev)
        net.latent_avg = to(dev, synth.randn("net3.latent_avg", (18, 512), seed, 0.1))
        img_in = synth.smooth_image("net3.img", 1, 512, seed)
        mask = synth.onehot(T(g["labels"].astype(np.int64)), 12)
        vec, struct = net.get_style_vectors(*to(dev, img_in, mask))
        codes = net.cal_style_codes(vec)
        out, inter = net(*to(dev, img_in, mask), randomize_noise=False)
        img2, minus1, _ = net.gen_img(struct, codes, to(dev, mask), randomize_noise=False)
    assert maxdiff(vec, g["vectors"]) < 1e-3
    assert codes.shape == (1, 12, 18, 512) and maxdiff(codes[:, :, :6], g["codes"]) < 1e-3
    assert maxdiff(out, g["image"]) < 1e-3
    assert minus1 == -1 and maxdiff(img2, out) == 0.0


@pytest.mark.parametrize("dev", DEVS)
def test_bisenet(golden, dev):
    from e4s2024_b200.face_parsing.model import BiSeNet
    g = golden("bisenet")
    with ctx_for(dev):
        seg = load(BiSeNet(19), 10, dev).eval()
        o, o16, o32 = seg(to(dev, T(g["x"])))
    assert o.shape == (2, 19, 128, 128)
    # synthetic weights give logits of magnitude ~2e2 -> relative tolerance (fp32 reassociation noise)
    for got, want in ((o[:, :, ::2, ::2], g["out"]), (o16[:, :, ::4, ::4], g["out16"]), (o32[:, :, ::4, ::4], g["out32"])):
        assert maxdiff(got, want) < 3e-5 * float(np.abs(want).max())


@pytest.mark.parametrize("dev", DEVS)
def test_face_parser(golden, dev):
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    p = golden("parser")
    with ctx_for(dev):
        parser = FaceParser(seg_ckpt=None, size=1024, device="cpu" if dev == "emul" else "cuda")
        load(parser.seg, 10, dev)
        img01 = (synth.smooth_image("parser.img", 1, 1024, 11) + 1) / 2
        pre = parser.preprocess_tensor(to(dev, img01))
        assert maxdiff(pre[..., :3].permute(0, 3, 1, 2)[:, :, ::8, ::8], p["pre_sample"]) < 1e-5
        lab12 = parser.parse_batch(to(dev, img01)).cpu().numpy()[0]
        lab19 = parser.parse_batch(to(dev, img01), convert_to_seg12=False).cpu().numpy()[0]
    # bit-exact label maps, except pixels where the reference's own top-2 margin is below fp32 noise
    for got, want in ((lab19, p["labels19"]), (lab12, p["labels12"])):
        bad = got != want
        tie = 2e-5 * float(p["logit_absmax"])          # fp32 reassociation noise at the synthetic logit scale
        assert bad.sum() == 0 or (bad.sum() <= 8 and float(p["margin"][bad].max()) < tie), int(bad.sum())


@pytest.mark.parametrize("dev", DEVS)
def test_swap_comp_style_vector(golden, dev):
    """the recombination step between encoder and generator, bit-exact vs the reference's function,
    plus a batch whose samples differ in whether the source has a mouth region (per-sample rule, no host read-back)."""
    from e4s2024_b200.swap_face_mask import swap_comp_style_vector
    g = golden("swap_comp_style_vector")
    with ctx_for(dev):
        ts, ss = [], []
        for i in range(int(g["n"])):
            cfg = g[f"cfg{i}"]
            t, s_ = to(dev, T(g[f"t{i}"]), T(g[f"s{i}"]))
            y = swap_comp_style_vector(t, s_, [int(c) for c in cfg[1:]], belowFace_interpolation=bool(cfg[0]))
            assert maxdiff(y, g[f"y{i}"]) == 0.0
            ts.append(T(g[f"t{i}"]))
            ss.append(T(g[f"s{i}"]))
        tb, sb = to(dev, torch.cat(ts), torch.cat(ss))
        yb = swap_comp_style_vector(tb, sb, [1, 2, 9], belowFace_interpolation=True)
        assert maxdiff(yb, orc.swap_comp_style_vector(torch.cat(ts), torch.cat(ss), [1, 2, 9], True)) == 0.0


@pytest.mark.parametrize("dev", DEVS)
def test_tensor2im_batch(golden, dev):
    """tensor2im's arithmetic on the device, byte-exact vs the reference's PIL output."""
    from e4s2024_b200.utils.torch_utils import tensor2im_batch
    g = golden("tensor2im")
    with ctx_for(dev):
        y = tensor2im_batch(to(dev, T(g["x"])))
    assert y.dtype == torch.uint8 and tuple(y.shape) == g["y"].shape
    assert int((y.cpu().numpy() != g["y"]).sum()) == 0


@pytest.mark.parametrize("dev", DEVS)
def test_im2tensor(golden, dev):
    """TO_TENSOR / NORMALIZE on the device, the same floats as torchvision's (every byte value covered)."""
    from e4s2024_b200 import _lib as L
    g = golden("im2tensor")
    with ctx_for(dev):
        y01, yn = L.im2tensor(to(dev, T(g["x"])))
    assert maxdiff(y01, g["y01"]) == 0.0 and maxdiff(yn, g["ynorm"]) == 0.0


MORPH_CASES = {"ones5": dict(kernel=torch.ones(5, 5)),
               "cross3": dict(kernel=torch.tensor([[0., 1, 0], [1, 1, 1], [0, 1, 0]])),
               "even4x6": dict(kernel=torch.ones(4, 6)),
               "nonflat": dict(kernel=torch.ones(3, 3), structuring_element=torch.tensor([[0., 0.1, 0], [0.1, 0.3, 0.1], [0, 0.1, 0]])),
               "const": dict(kernel=torch.ones(3, 5), border_type="constant", border_value=0.5),
               "origin": dict(kernel=torch.ones(3, 3), origin=[0, 2])}


@pytest.mark.parametrize("dev", DEVS)
def test_morphology(golden, dev):
    """SURVEY 8f row 4 (first piece): dilation / erosion / opening of the paste-back masks, bit-exact vs the reference functions."""
    from e4s2024_b200.utils import morphology as M
    g = golden("morphology")
    with ctx_for(dev):
        x = to(dev, T(g["x"]))
        for name, kw in MORPH_CASES.items():
            kw = {k: (to(dev, v) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
            assert maxdiff(M.dilation(x, **kw), g[name + "_dil"]) == 0.0, name
            assert maxdiff(M.erosion(x, **kw), g[name + "_ero"]) == 0.0, name
        k = to(dev, torch.ones(3, 3))
        assert maxdiff(M.opening(x, k), orc.morphology(orc.morphology(T(g["x"]), torch.ones(3, 3), False), torch.ones(3, 3), True)) == 0.0


@pytest.mark.parametrize("dev", DEVS)
def test_ops_autograd(dev):
    """upfirdn2d / fused_leaky_relu are differentiable like the reference ops (op/upfirdn2d.py:17-139, op/fused_act.py:18-69):
    first-order gradients and a second-order term against torch autograd through the oracle's pure-torch restatement."""
    from e4s2024_b200.stylegan2.op import fused_leaky_relu, upfirdn2d
    k4 = torch.tensor([1., 3., 3., 1.])
    k4 = torch.outer(k4, k4) / 64
    cases = [dict(kernel=k4 * 4, up=2, down=1, pad=(2, 1)), dict(kernel=k4 * 4, up=1, down=1, pad=(1, 1)),
             dict(kernel=k4, up=1, down=2, pad=(1, 1)), dict(kernel=torch.tensor([[1., 2., 1.], [0., 1., 3.]]) / 8, up=2, down=3, pad=(3, 0))]
    with ctx_for(dev):
        for i, kw in enumerate(cases):
            x0 = synth.randn(f"ag.up.x{i}", (2, 3, 9, 11), 90 + i)
            xr = x0.clone().requires_grad_(True)
            yr = orc.upfirdn2d(xr, kw["kernel"], kw["up"], kw["down"], kw["pad"])
            gy = synth.randn(f"ag.up.g{i}", tuple(yr.shape), 95 + i)
            (gr,) = torch.autograd.grad(yr, xr, gy)
            v = synth.randn(f"ag.up.v{i}", tuple(x0.shape), 99 + i)
            x = to(dev, x0).requires_grad_(True)
            gyd = to(dev, gy).requires_grad_(True)
            y = upfirdn2d(x, to(dev, kw["kernel"]), up=kw["up"], down=kw["down"], pad=kw["pad"])
            assert maxdiff(y.detach(), yr.detach()) < 1e-5
            (g,) = torch.autograd.grad(y, x, gyd, create_graph=True)
            assert maxdiff(g.detach(), gr.detach()) < 1e-5, i
            # second order: d<g, v>/d(gy) = upfirdn2d(v) -- the gradient is linear in gy
            (gg,) = torch.autograd.grad(g, gyd, to(dev, v))
            assert maxdiff(gg, orc.upfirdn2d(v, kw["kernel"], kw["up"], kw["down"], kw["pad"])) < 1e-5, i
        # fused_leaky_relu: grad wrt input and bias, and the second-order term wrt the incoming gradient
        x0 = synth.randn("ag.flr.x", (2, 6, 5, 7), 110)
        b0 = synth.randn("ag.flr.b", (6,), 111, 0.5)
        gy0 = synth.randn("ag.flr.g", (2, 6, 5, 7), 112)
        xr, br = x0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        yr = orc.fused_leaky_relu(xr, br)
        gxr, gbr = torch.autograd.grad(yr, (xr, br), gy0)
        x, b, gy = to(dev, x0).requires_grad_(True), to(dev, b0).requires_grad_(True), to(dev, gy0).requires_grad_(True)
        y = fused_leaky_relu(x, b)
        gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
        assert maxdiff(gx.detach(), gxr) < 1e-6 and maxdiff(gb.detach(), gbr) < 1e-5
        vx, vb = to(dev, synth.randn("ag.flr.vx", (2, 6, 5, 7), 113)), to(dev, synth.randn("ag.flr.vb", (6,), 114))
        (gg,) = torch.autograd.grad((gx, gb), gy, (vx, vb))
        ref = torch.where(yr.detach() > 0, 1.0, 0.2) * 2 ** 0.5 * (vx.cpu() + vb.cpu().reshape(1, -1, 1, 1))
        assert maxdiff(gg, ref) < 1e-6



def test_pack_cache_invalidation():
    """Packed-weight caches are keyed on (data_ptr, _version): `.data` writes (Ranger / EMA in the reference's training code) need
    engine.invalidate_packs, load_state_dict invalidates by itself."""
    from e4s2024_b200 import engine as E
    from e4s2024_b200.stylegan2.model import Generator
    with cpu_emul.emulated():
        G = load(Generator(32, 512, 2, remaining_layer_idx=18), 3, "emul").requires_grad_(False)
        latent = synth.randn("inval.latent", (1, 2, G.n_latent, 512), 3)
        mask = synth.onehot(synth.blocky_labels(1, 2, 32, cells=4, seed=3), 2)
        run = lambda: G([latent], None, mask, input_is_latent=True, randomize_noise=False)[0]
        a = run()
        sd = {k: v.clone() for k, v in G.state_dict().items()}
        G.convs[1].conv.weight.data[0, :, :, 1, 1].mul_(-1.0)   # does not bump _version
        stale = run()
        E.invalidate_packs(G)
        b = run()
        assert maxdiff(stale, a) == 0.0 and maxdiff(b, a) > 1e-3
        G.load_state_dict(sd)                                   # the post hook drops the caches
        assert maxdiff(run(), a) == 0.0


@pytest.mark.parametrize("dev", DEVS)
def test_soft_erosion(golden, dev):
    """SoftErosion (cone-kernel depthwise conv, min-iterations, threshold, global-max normalisation) vs the reference's
    own module output; hard-mask bits may differ only where the softened value sits on the threshold."""
    from e4s2024_b200.utils.paste_back_tricks import SoftErosion
    g = golden("soft_erosion")
    for name in ("default", "k7_it3"):
        ks, it = [int(v) for v in g[name + "_cfg"]]
        thr = float(g[name + "_thr"])
        with ctx_for(dev):
            mod = SoftErosion(kernel_size=ks, threshold=thr, iterations=it)
            mod = mod.to("cpu" if dev == "emul" else "cuda")
            y, mk = mod(to(dev, T(g["x"])))
        assert mk.dtype == torch.bool and tuple(y.shape) == g[name + "_y"].shape
        bad = mk.cpu().numpy() != g[name + "_mask"]
        assert bad.sum() <= 4, int(bad.sum())
        assert float((y.cpu() - T(g[name + "_y"])).abs()[torch.from_numpy(~bad)].max()) < 1e-5


@pytest.mark.parametrize("dev", DEVS)
def test_laplacian_pyramid_blend(golden, dev):
    """the multi-band blend (cv2.pyrDown / pyrUp arithmetic) vs the reference function's output, for the uint8 target
    frame + float composite the pipelines pass, and for all-float inputs; batched."""
    from e4s2024_b200.multi_band_blending import Laplacian_Pyramid_Blending_with_mask
    g = golden("laplacian_blend")
    lv = int(g["levels"])
    hwc = lambda a: T(a).permute(2, 0, 1)[None].contiguous()
    with ctx_for(dev):
        A8, Bf, m = to(dev, hwc(g["A8"]), hwc(g["B"]), hwc(g["m"]))
        out = Laplacian_Pyramid_Blending_with_mask(A8, Bf, m, lv)
        out_f = Laplacian_Pyramid_Blending_with_mask(A8.float(), Bf, m[:, :1], lv)          # one mask plane broadcast over channels
        two = Laplacian_Pyramid_Blending_with_mask(torch.cat([A8, A8]), torch.cat([Bf, Bf]), torch.cat([m, m]), lv)
    assert maxdiff(out[0].permute(1, 2, 0), g["out_u8A"]) < 1e-3
    assert maxdiff(out_f[0].permute(1, 2, 0), g["out_fA"]) < 1e-3
    assert maxdiff(two[1], out[0]) == 0.0

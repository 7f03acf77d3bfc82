"""CPU: the oracle restatement (oracle/e4s_oracle.py) against the committed golden vectors,
which are outputs of the reference's own modules (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc

T = torch.from_numpy


def close(a, b, tol):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= tol, float((a - b).abs().max())


def shapes_of(mod):
    return {k: v.shape for k, v in mod.items()}


def test_upfirdn2d(golden):
    g = golden("upfirdn2d")
    x = T(g["x"])
    for name in ("blur", "up2", "down2", "crop", "up2down3"):
        up, down, p0, p1 = [int(v) for v in g[name + "_cfg"]]
        close(orc.upfirdn2d(x, T(g[name + "_kernel"]), up, down, (p0, p1)), g[name], 1e-6)


def test_fused_leaky_relu(golden):
    g = golden("fused_leaky_relu")
    close(orc.fused_leaky_relu(T(g["x"]), T(g["bias"])), g["y"], 1e-6)
    close(orc.fused_leaky_relu(T(g["x"]), T(g["bias"]), 0.1, 1.5), g["y_slope01"], 1e-6)


MODCONV_SHAPES = lambda k: {"weight": (1, 16, 8, k, k), "modulation.weight": (8, 512), "modulation.bias": (8,)}


@pytest.mark.parametrize("tag,k,demod,up", [("k3", 3, True, False), ("k3up", 3, True, True), ("k1nodemod", 1, False, False)])
def test_modconv(golden, tag, k, demod, up):
    g = golden("modconv_" + tag)
    sd = synth.fill_state_dict(MODCONV_SHAPES(k), seed=3)
    y = orc.modulated_conv2d(T(g["x"]), T(g["style"]), sd, "", demodulate=demod, upsample=up)
    close(y, g["y"], 1e-5)


def styled_shapes():
    s = {"conv." + k: v for k, v in MODCONV_SHAPES(3).items()}
    s.update({"noise.weight": (1,), "activate.bias": (16,)})
    return s


@pytest.mark.parametrize("tag,up,seed", [("same", False, 4), ("up", True, 4), ("softmask", False, 6)])
def test_styledconv(golden, tag, up, seed):
    g = golden("styledconv_" + tag)
    sd = synth.fill_state_dict(styled_shapes(), seed=seed)
    y = orc.styled_conv(T(g["x"]), T(g["style"]), T(g["mask"]), sd, "", upsample=up, mask_op=True, noise=T(g["noise"]))
    close(y, g["y"], 1e-5)


def test_torgb(golden):
    g = golden("torgb")
    sd = synth.fill_state_dict({"conv.weight": (1, 3, 8, 1, 1), "conv.modulation.weight": (8, 512),
                                "conv.modulation.bias": (8,), "bias": (1, 3, 1, 1)}, seed=5)
    y = orc.to_rgb(T(g["x"]), T(g["style"]), T(g["mask"]), T(g["skip"]), sd, "", mask_op=True)
    close(y, g["y"], 1e-5)


def test_generator_small(golden):
    from e4s2024_b200.stylegan2.model import generator_state_shapes
    for tag in ("g32_rl18", "g64_rl5"):
        g = golden("generator_" + tag)
        size, rl, split, K, seed = [int(v) for v in g["cfg"]]
        sd = synth.fill_state_dict(generator_state_shapes(size), seed=seed)
        lab = T(g["labels"].astype(np.int64))
        n_latent = int(np.log2(size)) * 2 - 2
        latent = synth.randn(f"{tag}.latent", (2, K, n_latent, 512), seed)
        img, inter = orc.generator_forward(sd, size, latent, synth.onehot(lab, K), split_layer_idx=split,
                                           remaining_layer_idx=rl)
        close(img, g["image"], 1e-4)
        close(inter[:, ::16], g["inter"], 1e-4)


def test_encoder(golden):
    from e4s2024_b200.encoders.psp_encoders import encoder_state_shapes
    g = golden("encoder")
    sd = synth.fill_state_dict(encoder_state_shapes(), seed=8)
    mask = synth.onehot(T(g["labels"].astype(np.int64)), 12)
    codes, struct = orc.fs_encoder_psp(sd, T(g["x"]), mask)
    close(codes, g["codes"], 1e-4)
    assert float(codes[1, 3].abs().max()) == 0.0          # empty region -> zero vector
    assert struct.shape == (2, 512, 8, 8) and float(struct.abs().max()) == 0.0


def test_bisenet_and_parser(golden):
    from e4s2024_b200.face_parsing.model import bisenet_state_shapes
    g = golden("bisenet")
    sd = synth.fill_state_dict(bisenet_state_shapes(19), seed=10)
    o, o16, o32 = orc.bisenet_forward(sd, T(g["x"]))
    close(o[:, :, ::2, ::2], g["out"], 1e-4)
    close(o16[:, :, ::4, ::4], g["out16"], 1e-4)
    close(o32[:, :, ::4, ::4], g["out32"], 1e-4)
    p = golden("parser")
    img01 = (synth.smooth_image("parser.img", 1, 1024, 11) + 1) / 2
    close(orc.parser_preprocess(img01, 1024)[:, :, ::8, ::8], p["pre_sample"], 1e-6)
    lab = orc.face_parse(sd, img01)[0]
    bad = lab != p["labels12"]
    assert bad.sum() == 0


def test_seg_lut(golden):
    g = golden("seg19_to_seg12")
    assert (orc.SEG19_TO_SEG12[g["src"]] == g["dst"]).all()
    lab = torch.tensor([[[[0, 2], [1, 1]]]])
    oh = orc.label_to_onehot(lab, 3)
    assert oh.shape == (1, 3, 2, 2) and oh.sum() == 4 and oh[0, 2, 0, 1] == 1


def test_swap_comp_style_vector(golden):
    g = golden("swap_comp_style_vector")
    for i in range(int(g["n"])):
        cfg = g[f"cfg{i}"]
        y = orc.swap_comp_style_vector(T(g[f"t{i}"]), T(g[f"s{i}"]), [int(c) for c in cfg[1:]], bool(cfg[0]))
        close(y, g[f"y{i}"], 0.0)


def test_im2tensor(golden):
    g = golden("im2tensor")
    x01, xn = orc.to_tensor_normalize(T(g["x"]))
    assert float((x01 - T(g["y01"])).abs().max()) == 0.0 and float((xn - T(g["ynorm"])).abs().max()) == 0.0


def test_tensor2im(golden):
    g = golden("tensor2im")
    assert int((orc.tensor2im_u8(T(g["x"])).numpy() != g["y"]).sum()) == 0


MORPH_CASES = {"ones5": dict(kernel=torch.ones(5, 5)),
               "cross3": dict(kernel=torch.tensor([[0., 1, 0], [1, 1, 1], [0, 1, 0]])),
               "even4x6": dict(kernel=torch.ones(4, 6)),
               "nonflat": dict(kernel=torch.ones(3, 3), structuring_element=torch.tensor([[0., 0.1, 0], [0.1, 0.3, 0.1], [0, 0.1, 0]])),
               "const": dict(kernel=torch.ones(3, 5), border_type="constant", border_value=0.5),
               "origin": dict(kernel=torch.ones(3, 3), origin=[0, 2])}


def test_morphology(golden):
    g = golden("morphology")
    x = T(g["x"])
    for name, kw in MORPH_CASES.items():
        okw = {k: v for k, v in kw.items() if k != "kernel"}
        close(orc.morphology(x, kw["kernel"], True, **okw), g[name + "_dil"], 0.0)
        close(orc.morphology(x, kw["kernel"], False, **okw), g[name + "_ero"], 0.0)



def test_soft_erosion(golden):
    g = golden("soft_erosion")
    for name in ("default", "k7_it3"):
        ks, it = [int(v) for v in g[name + "_cfg"]]
        y, mk = orc.soft_erosion(T(g["x"]), kernel_size=ks, threshold=float(g[name + "_thr"]), iterations=it)
        assert float((y - T(g[name + "_y"])).abs().max()) < 1e-6 and int((mk.numpy() != g[name + "_mask"]).sum()) == 0


def test_laplacian_blend(golden):
    g = golden("laplacian_blend")
    lv = int(g["levels"])
    out = orc.laplacian_pyramid_blend(g["A8"], g["B"], g["m"], lv)
    assert float(np.abs(out - g["out_u8A"]).max()) < 1e-3                  # [0,255] scale; cv2's SIMD sums in another order
    out = orc.laplacian_pyramid_blend(g["A8"].astype(np.float32), g["B"], g["m"], lv)
    assert float(np.abs(out - g["out_fA"]).max()) < 1e-3

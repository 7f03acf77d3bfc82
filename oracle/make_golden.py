"""Generate tests/golden/*.npz by executing the REFERENCE's own modules (CPU, fp32).

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Each fixture stores the small inputs and the reference outputs; weights are NOT stored, they
are regenerated from e4s2024_b200.synth (deterministic numpy PCG64 streams keyed by name).
The script also prints the oracle-vs-reference difference for every case, which is how the
oracle restatement (oracle/e4s_oracle.py) was pinned.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from e4s2024_b200 import synth  # noqa: E402
from oracle import e4s_oracle as orc  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})


def diff(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    ref_shims.install()
    sg = ref_shims.stylegan_module()
    report = {}

    # ---- upfirdn2d: the two generator modes + a down-sampling and a crop case -------------------
    from models.stylegan2.op import upfirdn2d as ref_upfirdn2d, fused_leaky_relu as ref_flr
    x = synth.randn("upfirdn.x", (2, 3, 9, 11), 1)
    k4 = sg.make_kernel([1, 3, 3, 1])
    cases = {"blur": dict(kernel=k4 * 4, up=1, down=1, pad=(1, 1)),
             "up2": dict(kernel=k4 * 4, up=2, down=1, pad=(2, 1)),
             "down2": dict(kernel=k4, up=1, down=2, pad=(1, 1)),
             "crop": dict(kernel=sg.make_kernel([1, 2, 1]), up=1, down=1, pad=(-1, 2)),
             "up2down3": dict(kernel=k4, up=2, down=3, pad=(3, 0))}
    arrs = {"x": x.numpy()}
    for name, kw in cases.items():
        y = ref_upfirdn2d(x, **kw)
        report["upfirdn2d/" + name] = diff(y, orc.upfirdn2d(x, **kw))
        arrs[name] = y.numpy()
        arrs[name + "_kernel"] = kw["kernel"].numpy()
        arrs[name + "_cfg"] = np.array([kw["up"], kw["down"], kw["pad"][0], kw["pad"][1]])
    save("upfirdn2d", **arrs)

    # ---- fused_leaky_relu (formula of the .cu; there is no CPU code in the reference) -----------
    xa = synth.randn("flr.x", (2, 6, 5, 7), 1)
    ba = synth.randn("flr.b", (6,), 1)
    save("fused_leaky_relu", x=xa.numpy(), bias=ba.numpy(), y=ref_flr(xa, ba).numpy(),
         y_slope01=ref_flr(xa, ba, 0.1, 1.5).numpy())

    # ---- ModulatedConv2d / StyledConv / ToRGB on small channel counts ----------------------------
    for tag, kw in {"k3": dict(kernel_size=3), "k3up": dict(kernel_size=3, upsample=True),
                    "k1nodemod": dict(kernel_size=1, demodulate=False)}.items():
        m = sg.ModulatedConv2d(8, 16, style_dim=512, **kw)
        sd = synth.synth_module_weights(m, seed=3)
        xi = synth.randn(f"modconv.{tag}.x", (2, 8, 6, 6), 3)
        st = synth.randn(f"modconv.{tag}.s", (2, 512), 3)
        y = m(xi, st)
        yo = orc.modulated_conv2d(xi, st, sd, "", demodulate=kw.get("demodulate", True),
                                  upsample=kw.get("upsample", False))
        report["modconv/" + tag] = diff(y, yo)
        save("modconv_" + tag, x=xi.numpy(), style=st.numpy(), y=y.numpy())

    K = 5
    lab = synth.blocky_labels(2, K, 16, cells=4, seed=4)
    mask = synth.onehot(lab, K)
    for tag, up in (("same", False), ("up", True)):
        m = sg.StyledConv(8, 16, 3, 512, upsample=up, mask_op=True)
        sd = synth.synth_module_weights(m, seed=4)
        xi = synth.randn(f"styled.{tag}.x", (2, 8, 8, 8), 4)
        st = synth.randn(f"styled.{tag}.s", (2, K, 512), 4)
        r = 16 if up else 8
        nz = synth.randn(f"styled.{tag}.n", (1, 1, r, r), 4)
        y = m(xi, st, mask, noise=nz)
        yo = orc.styled_conv(xi, st, mask, sd, "", upsample=up, mask_op=True, noise=nz)
        report["styled/" + tag] = diff(y, yo)
        save("styledconv_" + tag, x=xi.numpy(), style=st.numpy(), mask=mask.numpy(), noise=nz.numpy(), y=y.numpy())
    m = sg.ToRGB(8, 512, upsample=True, mask_op=True)
    sd = synth.synth_module_weights(m, seed=5)
    xi = synth.randn("torgb.x", (2, 8, 8, 8), 5)
    st = synth.randn("torgb.s", (2, K, 512), 5)
    sk = synth.randn("torgb.skip", (2, 3, 4, 4), 5)
    y = m(xi, st, mask, sk)
    report["torgb"] = diff(y, orc.to_rgb(xi, st, mask, sk, sd, "", mask_op=True))
    save("torgb", x=xi.numpy(), style=st.numpy(), mask=mask.numpy(), skip=sk.numpy(), y=y.numpy())

    # non-one-hot (soft) masks are legal inputs
    soft = torch.softmax(synth.randn("soft.mask", (2, K, 16, 16), 6), dim=1)
    m = sg.StyledConv(8, 16, 3, 512, upsample=False, mask_op=True)
    sd = synth.synth_module_weights(m, seed=6)
    xi = synth.randn("soft.x", (2, 8, 8, 8), 6)
    st = synth.randn("soft.s", (2, K, 512), 6)
    nz = synth.randn("soft.n", (1, 1, 8, 8), 6)
    y = m(xi, st, soft, noise=nz)
    report["styled/softmask"] = diff(y, orc.styled_conv(xi, st, soft, sd, "", upsample=False, mask_op=True, noise=nz))
    save("styledconv_softmask", x=xi.numpy(), style=st.numpy(), mask=soft.numpy(), noise=nz.numpy(), y=y.numpy())

    # ---- full Generator: 32^2 all-masked (default rl) and 64^2 with un-masked tail (rl=5) --------
    for tag, size, rl, split, K in (("g32_rl18", 32, 18, 7, 12), ("g64_rl5", 64, 5, 5, 4)):
        G = ref_shims.generator_cls()(size, 512, 8, split_layer_idx=split, remaining_layer_idx=rl).eval()
        sd = synth.synth_module_weights(G, seed=7)
        lab = synth.blocky_labels(2, K, size, cells=8, seed=7)
        mask = synth.onehot(lab, K)
        latent = synth.randn(f"{tag}.latent", (2, K, G.n_latent, 512), 7)
        img, _, inter = G([latent], None, mask, input_is_latent=True, randomize_noise=False)
        io, into = orc.generator_forward(sd, size, latent, mask, split_layer_idx=split, remaining_layer_idx=rl)
        report["generator/" + tag] = diff(img, io)
        report["generator/" + tag + "/inter"] = diff(inter, into)
        save("generator_" + tag, labels=lab.numpy().astype(np.uint8), image=img.numpy(), inter=inter.numpy()[:, ::16],
             cfg=np.array([size, rl, split, K, 7]))

    # ---- FSEncoder_PSP on a 128^2 input ----------------------------------------------------------
    enc = ref_shims.encoder_cls()(mode="ir_se", opts=None).eval()
    sd = synth.synth_module_weights(enc, seed=8)
    xi = synth.smooth_image("enc.x", 2, 128, 8)
    lab = synth.blocky_labels(2, 12, 64, cells=8, seed=8)
    lab[1][lab[1] == 3] = 0                               # sample 1 has an empty region -> zero vector
    mask = synth.onehot(lab, 12)
    codes, struct = enc(xi, mask)
    co, _ = orc.fs_encoder_psp(sd, xi, mask)
    report["encoder/codes"] = diff(codes, co)
    save("encoder", x=xi.numpy(), labels=lab.numpy().astype(np.uint8), codes=codes.numpy())

    # ---- Net3.forward at out_size 64 (rl=5 -> both masked and un-masked layers) -------------------
    net = ref_shims.net3_cls()(ref_shims.net3_opts(out_size=64, remaining_layer_idx=5)).eval()
    sd = synth.synth_module_weights(net, seed=9)
    net.latent_avg = synth.randn("net3.latent_avg", (18, 512), 9, 0.1)
    img_in = synth.smooth_image("net3.img", 1, 512, 9)
    lab = synth.blocky_labels(1, 12, 64, cells=8, seed=9)
    mask = synth.onehot(lab, 12)
    out, inter = net(img_in, mask, randomize_noise=False)
    oo, oi, ocodes, ovec = orc.net3_forward(sd, img_in, mask, net.latent_avg, out_size=64, remaining_layer_idx=5)
    report["net3/image"] = diff(out, oo)
    vec, _ = net.get_style_vectors(img_in, mask)
    codes = net.cal_style_codes(vec)
    report["net3/codes"] = diff(codes, ocodes)
    save("net3", labels=lab.numpy().astype(np.uint8), image=out.numpy(), vectors=vec.numpy(),
         codes=codes.numpy()[:, :, :6], cfg=np.array([64, 5, 9]))

    # ---- BiSeNet logits at 128^2 + full parser front-end (bicubic 1024->512, argmax, LUT) ---------
    bm = ref_shims.bisenet_module()
    seg = bm.BiSeNet(n_classes=19).eval()
    sd = synth.synth_module_weights(seg, seed=10)
    xi = synth.randn("bisenet.x", (2, 3, 128, 128), 10)
    o, o16, o32 = seg(xi)
    q, q16, q32 = orc.bisenet_forward(sd, xi)
    report["bisenet/out"] = max(diff(o, q), diff(o16, q16), diff(o32, q32))
    save("bisenet", x=xi.numpy(), out=o.numpy()[:, :, ::2, ::2], out16=o16.numpy()[:, :, ::4, ::4], out32=o32.numpy()[:, :, ::4, ::4])

    down = ref_shims.bicubic_cls()(factor=2)
    down.cuda = ""
    img01 = (synth.smooth_image("parser.img", 1, 1024, 11) + 1) / 2
    pre = (down(img01).clamp(0, 1) - torch.tensor(orc.SEG_MEAN).view(1, 3, 1, 1)) / torch.tensor(orc.SEG_STD).view(1, 3, 1, 1)
    report["parser/preprocess"] = diff(pre, orc.parser_preprocess(img01, 1024))
    logits = seg(pre)[0]
    lab19 = torch.argmax(logits, dim=1)[0].numpy().astype(np.uint8)
    lab12 = ref_shims.seg19_to_seg12_fn()(lab19)
    top2 = torch.topk(logits, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[0].numpy()
    report["parser/labels_mismatch"] = float((orc.face_parse(sd, img01)[0] != lab12).sum())
    save("parser", pre_sample=pre.numpy()[:, :, ::8, ::8], labels19=lab19, labels12=lab12, margin=margin.astype(np.float32),
         logit_absmax=np.float32(logits.abs().max()))
    d4 = ref_shims.bicubic_cls()(factor=4)
    d4.cuda = ""
    report["parser/bicubic4"] = diff(d4(img01[:, :, :256, :256]), orc.bicubic_downsample(img01[:, :, :256, :256], 4))

    lut_in = np.arange(19, dtype=np.uint8).reshape(1, 19)
    save("seg19_to_seg12", src=lut_in, dst=ref_shims.seg19_to_seg12_fn()(lut_in))
    report["lut"] = float((orc.SEG19_TO_SEG12[lut_in] != ref_shims.seg19_to_seg12_fn()(lut_in)).sum())

    print("\noracle vs reference (max |diff|):")
    for k, v in report.items():
        print(f"  {k:32s} {v:.3e}")
    with open(os.path.join(OUT, "PINNING_REPORT.txt"), "w") as f:
        f.write("oracle/e4s_oracle.py vs reference modules executed from /root/reference (CPU fp32), max |diff|\n")
        for k, v in report.items():
            f.write(f"{k:32s} {v:.3e}\n")


if __name__ == "__main__":
    main()

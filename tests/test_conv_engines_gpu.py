"""GPU: the tcgen05 (bf16x3 split) convolution engine against the exact-fp32 CUDA-core engine and a plain
torch fp64 reference of the same op, over the geometries the hot path uses.  Tolerance: the split keeps
~16 mantissa bits per operand -> relative error ~1e-5 of the output scale (stated bound 2e-4)."""
import pytest
import torch
import torch.nn.functional as F

from e4s2024_b200 import synth

pytestmark = pytest.mark.gpu


def _mk(shape, name, std=1.0):
    return synth.randn(name, shape, 21, std).cuda()


CASES = [
    # (name, B, H, W, Cin, Cout, k, stride, up2, regions)
    ("3x3_c64_n64", 2, 16, 16, 64, 64, 3, 1, False, 1),
    ("3x3_c128_n32", 1, 24, 20, 128, 32, 3, 1, False, 1),
    ("3x3_c512_n512_regions", 2, 8, 8, 512, 512, 3, 1, False, 5),
    ("3x3_s2_c64_n128", 2, 18, 18, 64, 128, 3, 2, False, 1),
    ("1x1_c256_n256", 2, 8, 8, 256, 256, 1, 1, False, 1),
    ("up2_c128_n64_regions", 2, 8, 8, 128, 64, 3, 1, True, 4),
    ("tiny_m", 1, 4, 4, 512, 512, 3, 1, False, 3),
    ("3x3_c32_n32", 1, 40, 40, 32, 32, 3, 1, False, 1),
    ("3x3_c8_n64_rgb", 2, 20, 20, 8, 64, 3, 1, False, 1),
    ("up2_c64_n32", 1, 12, 12, 64, 32, 3, 1, True, 1),
    ("3x3_c16_n128_s2", 1, 14, 14, 16, 128, 3, 2, False, 2),
    # geometries taken by the halo kernel (H % 16 == 0, W % 8 == 0, one style per sample)
    ("halo_c32_n32", 1, 32, 48, 32, 32, 3, 1, False, 1),
    ("halo_c64_n64", 3, 32, 24, 64, 64, 3, 1, False, 1),
    ("halo_c128_n256", 1, 32, 16, 128, 256, 3, 1, False, 1),
    ("halo_c512_n512", 1, 16, 16, 512, 512, 3, 1, False, 1),
    ("halo_up2_c64_n32", 2, 16, 16, 64, 32, 3, 1, True, 1),
    ("halo_up2_c128_n64", 2, 16, 8, 128, 64, 3, 1, True, 1),
    ("halo_up2_c32_n32", 1, 32, 16, 32, 32, 3, 1, True, 1),
    # halo kernel with a full persistent grid (>= 148 jobs, uneven job counts per CTA): cluster launch, weights multicast
    ("halo_cluster_c64_n64", 2, 128, 128, 64, 64, 3, 1, False, 1),
    ("halo_cluster_up2_c64_n32", 2, 128, 128, 64, 32, 3, 1, True, 1),
    ("halo_cluster_up2_c128_n128", 1, 128, 160, 128, 128, 3, 1, True, 1),
    # geometries taken by the 512-column gather kernel (conv_tc_wide.cu): merged tiles (one region per sample or a coarse
    # mask), mixed tiles (per-pixel random labels -> the four phases of a pixel differ -> per-phase passes)
    ("wide_up2_c128_n128", 2, 16, 16, 128, 128, 3, 1, True, 1),
    ("wide_up2_c64_n256_mixed", 1, 32, 16, 64, 256, 3, 1, True, 6),
    ("wide_up2_c64_n128_coarse", 2, 16, 16, 64, 128, 3, 1, True, -4),
    ("wide_c64_n512_regions", 1, 32, 16, 64, 512, 3, 1, False, 7),
    ("wide_c128_n1024", 1, 16, 8, 128, 1024, 3, 1, False, 1),
    # stride-2 3x3 (the encoder's down-sampling convs) on the wide kernel: taken from 100 output tiles up (4 x 4 x 8 = 128 here)
    ("wide_s2_c64_n512", 4, 128, 128, 64, 512, 3, 2, False, 1),
    ("wide_s2_c64_n512_regions", 4, 128, 128, 64, 512, 3, 2, False, 3),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_matches_f32_and_torch(case):
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    if not E.tc_available():
        pytest.skip("library built without the tcgen05 engine")
    name, B, H, W, Cin, Cout, k, stride, up2, R = case
    coarse = R < 0            # labels constant over 8x8 blocks of the 32x32 label map: phase-uniform rows (merged tiles)
    R = abs(R)
    x = _mk((B, H, W, Cin), name + ".x")
    w = _mk((Cout, Cin, k, k), name + ".w", (1.0 / (Cin * k * k)) ** 0.5)
    Ho, Wo = (2 * H, 2 * W) if up2 else ((H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1)
    smod = (1.0 + 0.3 * _mk((B, R, Cin), name + ".s")).contiguous()
    demod = (1.0 + 0.2 * _mk((B, R, Cout), name + ".d")).contiguous()
    labels = torch.randint(0, R, (B, 32, 32), device="cuda", dtype=torch.uint8) if R > 1 else None
    if coarse:
        labels = torch.randint(0, R, (B, 4, 4), device="cuda", dtype=torch.uint8).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
    noise = _mk((1, 1, Ho, Wo), name + ".n")
    nw = torch.tensor([0.1], device="cuda")
    bias = _mk((Cout,), name + ".b", 0.1)
    if up2:
        fir = torch.tensor([1., 3., 3., 1.])
        fir = (torch.outer(fir, fir) / 64 * 4).cuda()
        pw = E.pack_up_weight(w, fir)
    else:
        pw = E.pack_conv_weight(w)
    assert pw.tc is not None
    kw = dict(stride=stride, up2=up2, smod=smod, demod=demod, labels=labels, regions=R, noise=noise, noise_w=nw, ch_shift=bias,
              act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
    y32 = E.conv(E.View(x), pw, engine="f32", **kw).t
    ytc = E.conv(E.View(x), pw, engine="tc", **kw).t
    torch.cuda.synchronize()
    scale = float(y32.abs().max())
    d = float((ytc - y32).abs().max())
    print(f"{name}: max|tc - f32| = {d:.3e} (output scale {scale:.2f})")
    assert d < 2e-4 * max(scale, 1.0)

    # torch fp64 reference of the same op (per-pixel region modulation done the slow way)
    xd, wd = x.double().permute(0, 3, 1, 2), w.double()
    ys = torch.arange(Ho, device="cuda") * 32 // Ho
    xs = torch.arange(Wo, device="cuda") * 32 // Wo
    reg = labels.long()[:, ys][:, :, xs] if labels is not None else torch.zeros(B, Ho, Wo, dtype=torch.long, device="cuda")
    ref = torch.zeros(B, Cout, Ho, Wo, dtype=torch.float64, device="cuda")
    for r in range(R):
        xm = xd * smod[:, r].double()[:, :, None, None]
        if up2:
            yt = F.conv_transpose2d(xm, wd.transpose(0, 1), stride=2)
            kf = fir.double()[None, None].repeat(Cout, 1, 1, 1)
            yr = F.conv2d(F.pad(yt, (1, 1, 1, 1)), torch.flip(kf, [2, 3]), groups=Cout)
        else:
            yr = F.conv2d(xm, wd, stride=stride, padding=k // 2)
        yr = yr * demod[:, r].double()[:, :, None, None]
        ref += yr * (reg == r)[:, None].double()
    ref = ref + 0.1 * noise.double() + bias.double()[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) * 2 ** 0.5
    d32 = float((y32.permute(0, 3, 1, 2).double() - ref).abs().max())
    dtc = float((ytc.permute(0, 3, 1, 2).double() - ref).abs().max())
    print(f"{name}: vs torch fp64: f32 engine {d32:.3e}, tc engine {dtc:.3e}")
    assert d32 < 2e-5 * max(scale, 1.0) and dtc < 2e-4 * max(scale, 1.0)


def test_weight_pack_kernels_match_the_torch_derivation():
    """e4s_pack_conv_weights_f32 / e4s_pack_upconv_weights_f32 (one launch per layer) against the torch restatement in
    tests/cpu_emul.py (the poly-phase derivation of SURVEY.md appendix B.1, itself checked against the reference's
    conv_transpose2d + Blur by the modconv goldens)."""
    from e4s2024_b200 import _lib as L
    from tests import cpu_emul
    g = torch.Generator().manual_seed(3)
    for co, ci, k, cin_pad in ((19, 24, 1, 24), (64, 3, 7, 8), (32, 16, 3, 16)):
        w = torch.randn(co, ci, k, k, generator=g)
        cout_pad = (co + 3) // 4 * 4
        got = L.pack_conv_weights(w.cuda(), cin_pad, cout_pad, 0.37).cpu()
        assert torch.equal(got, cpu_emul.pack_conv_weights(w, cin_pad, cout_pad, 0.37))
        got = L.pack_conv_weights(w.cuda(), cin_pad, cout_pad, 0.37, True).cpu()
        want = cpu_emul.pack_conv_weights(w, cin_pad, cout_pad, 0.37, True)
        assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max())
    w = torch.randn(12, 16, 3, 3, generator=g)
    k1 = torch.tensor([1., 3., 3., 1.])
    fir = torch.outer(k1, k1) / 64 * 4
    got = L.pack_upconv_weights(w.cuda(), fir.cuda(), 12, 0.25).cpu()
    want = cpu_emul.pack_upconv_weights(w, fir, 12, 0.25)
    assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max())


UPZ_CASES = [
    # (name, B, H, W, Cin, Cout, regions, label kind): "blocky" = 8x8 cells of the 32x32 label map, "pixel" = independent per label pixel
    ("upz_c64_n128_blocky", 2, 16, 16, 64, 128, 4, "blocky"),
    ("upz_c128_n256_pixel", 1, 16, 8, 128, 256, 6, "pixel"),
    ("upz_c64_n512_odd_blocky", 2, 12, 20, 64, 512, 5, "blocky"),
    ("upz_c512_n512_tiny", 3, 4, 4, 512, 512, 12, "pixel"),
    ("upz_c256_n128_one_region", 1, 24, 24, 256, 128, 3, "one"),
]


def _upz_rows(labels, B, H, W, ratio=32.0, lazy=False):
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    cells_total = B * (H + 1) * (W + 1)
    max_rows = E.pad_to(int(ratio * cells_total), 128)
    cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    cells, rows = L.upz_build_rows(labels, H, W, max_rows, cnt)
    return E.UpzRows(cells, rows, cnt, None if lazy else int(cnt.item()), max_rows, max_rows, cells_total)


@pytest.mark.parametrize("case", UPZ_CASES, ids=[c[0] for c in UPZ_CASES])
def test_upz_matches_polyphase_and_torch(case):
    """Masked up-convolution as the (cell, region) conv_transpose GEMM + FIR pass (csrc/conv_tc_upz.cu) against the exact-fp32 poly-phase
    engine and torch fp64 (conv_transpose2d + blur per region, masked after the blur: reference model.py:287-300,395-398)."""
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    if not E.tc_available():
        pytest.skip("library built without the tcgen05 engine")
    name, B, H, W, Cin, Cout, R, kind = case
    x = _mk((B, H, W, Cin), name + ".x")
    w = _mk((Cout, Cin, 3, 3), name + ".w", (1.0 / (Cin * 9)) ** 0.5)
    Ho, Wo = 2 * H, 2 * W
    smod = (1.0 + 0.3 * _mk((B, R, Cin), name + ".s")).contiguous()
    demod = (1.0 + 0.2 * _mk((B, R, Cout), name + ".d")).contiguous()
    g = torch.Generator(device="cuda").manual_seed(5)
    if kind == "blocky":
        labels = torch.randint(0, R, (B, 4, 4), device="cuda", dtype=torch.uint8, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
    elif kind == "pixel":
        labels = torch.randint(0, R, (B, 32, 32), device="cuda", dtype=torch.uint8, generator=g)
    else:
        labels = torch.full((B, 32, 32), R - 1, device="cuda", dtype=torch.uint8)
    noise = _mk((1, 1, Ho, Wo), name + ".n")
    nw = torch.tensor([0.1], device="cuda")
    bias = _mk((Cout,), name + ".b", 0.1)
    fir = torch.tensor([1., 3., 3., 1.])
    fir = (torch.outer(fir, fir) / 64 * 4).cuda()
    pw = E.pack_up_weight(w, fir)
    assert pw.tc is not None and pw.tc9 is not None
    kw = dict(up2=True, smod=smod, demod=demod, labels=labels, regions=R, noise=noise, noise_w=nw, ch_shift=bias,
              act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
    y32 = E.conv(E.View(x), pw, engine="f32", **kw).t
    uz = _upz_rows(labels, B, H, W)
    launches = L.launch_count()
    yz = E.conv(E.View(x), pw, engine="tc", upz=uz, **kw).t
    assert L.launch_count() - launches == 2                     # the GEMM and the FIR pass, nothing else
    # lazy context (count stays on the device): limit not exceeded -> same kernels, same bits; limit exceeded -> the poly-phase kernel
    uz_lazy = _upz_rows(labels, B, H, W, lazy=True)
    yz_lazy = E.conv(E.View(x), pw, engine="tc", upz=uz_lazy, **kw).t
    uz_small = _upz_rows(labels, B, H, W, lazy=True)
    uz_small.limit = 0
    ypoly = E.conv(E.View(x), pw, engine="tc", upz=uz_small, **kw).t
    yplain = E.conv(E.View(x), pw, engine="tc", **kw).t
    torch.cuda.synchronize()
    scale = float(y32.abs().max())
    d = float((yz - y32).abs().max())
    rows_per_cell = uz.count / uz.cells_total
    print(f"{name}: {rows_per_cell:.2f} rows per cell, max|upz - f32| = {d:.3e} (output scale {scale:.2f})")
    assert d < 2e-4 * max(scale, 1.0)
    assert torch.equal(yz, yz_lazy)
    assert torch.equal(ypoly, yplain)
    if kind == "one":
        assert uz.count == uz.cells_total

    xd, wd = x.double().permute(0, 3, 1, 2), w.double()
    ys = torch.arange(Ho, device="cuda") * 32 // Ho
    xs = torch.arange(Wo, device="cuda") * 32 // Wo
    reg = labels.long()[:, ys][:, :, xs]
    ref = torch.zeros(B, Cout, Ho, Wo, dtype=torch.float64, device="cuda")
    for r in range(R):
        xm = xd * smod[:, r].double()[:, :, None, None]
        yt = F.conv_transpose2d(xm, wd.transpose(0, 1), stride=2)
        kf = fir.double()[None, None].repeat(Cout, 1, 1, 1)
        yr = F.conv2d(F.pad(yt, (1, 1, 1, 1)), torch.flip(kf, [2, 3]), groups=Cout)
        yr = yr * demod[:, r].double()[:, :, None, None]
        ref += yr * (reg == r)[:, None].double()
    ref = ref + 0.1 * noise.double() + bias.double()[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) * 2 ** 0.5
    dz = float((yz.permute(0, 3, 1, 2).double() - ref).abs().max())
    print(f"{name}: vs torch fp64: upz {dz:.3e}")
    assert dz < 2e-4 * max(scale, 1.0)


def test_upz_row_list_is_the_set_of_regions_reading_each_cell():
    """e4s_upz_build_rows against a torch restatement: cell (cy,cx) lists region r iff one of the output pixels (2cy-2..2cy+2, 2cx-2..2cx+2)
    lies in r; rows of a cell are consecutive, ascending in r, and `count` is the total."""
    B, H, W, R = 2, 8, 12, 7
    g = torch.Generator(device="cuda").manual_seed(11)
    labels = torch.randint(0, R, (B, 32, 48), device="cuda", dtype=torch.uint8, generator=g)
    uz = _upz_rows(labels, B, H, W)
    cells, rows = uz.cells.cpu(), uz.rows.cpu()
    lab = labels.cpu().long()
    ys = torch.arange(2 * H) * 32 // (2 * H)
    xs = torch.arange(2 * W) * 48 // (2 * W)
    reg = lab[:, ys][:, :, xs]
    total = 0
    for b in range(B):
        for cy in range(H + 1):
            for cx in range(W + 1):
                win = reg[b, max(2 * cy - 2, 0):min(2 * cy + 2, 2 * H - 1) + 1, max(2 * cx - 2, 0):min(2 * cx + 2, 2 * W - 1) + 1]
                want = sorted(set(win.reshape(-1).tolist()))
                mask, base = int(cells[b, cy, cx, 0]), int(cells[b, cy, cx, 1])
                assert mask == sum(1 << r for r in want)
                for j, r in enumerate(want):
                    assert int(rows[base + j, 0]) == (b << 8 | r) and int(rows[base + j, 1]) == (cy << 16 | cx)
                total += len(want)
    assert total == uz.count


@pytest.mark.parametrize("cin,cout,H,W", [(64, 32, 20, 12), (128, 64, 16, 16), (64, 128, 8, 24)])
def test_upz_direct_form_matches_the_polyphase_kernels(cin, cout, H, W, monkeypatch):
    """The direct (un-masked) form of e4s_conv_tc_upz -- opt-in, E4S_UPZ_DIRECT=1: one row per cell, slot widths 32 / 64 / 128 -- against the
    exact-fp32 poly-phase engine.  (It is off by default because it is slower on B200, see engine.upz_eligible.)"""
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    if not E.tc_available():
        pytest.skip("library built without the tcgen05 engine")
    monkeypatch.setenv("E4S_UPZ_DIRECT", "1")
    name = f"upzd_{cin}_{cout}"
    B = 2
    x = _mk((B, H, W, cin), name + ".x")
    w = _mk((cout, cin, 3, 3), name + ".w", (1.0 / (cin * 9)) ** 0.5)
    smod = (1.0 + 0.3 * _mk((B, 1, cin), name + ".s")).contiguous()
    demod = (1.0 + 0.2 * _mk((B, 1, cout), name + ".d")).contiguous()
    noise = _mk((1, 1, 2 * H, 2 * W), name + ".n")
    nw = torch.tensor([0.1], device="cuda")
    bias = _mk((cout,), name + ".b", 0.1)
    fir = torch.tensor([1., 3., 3., 1.])
    fir = (torch.outer(fir, fir) / 64 * 4).cuda()
    pw = E.pack_up_weight(w, fir)
    assert pw.tc9 is not None
    kw = dict(up2=True, smod=smod, demod=demod, regions=1, noise=noise, noise_w=nw, ch_shift=bias, act=L.ACT_LRELU, slope=0.2, gain=2 ** 0.5)
    y32 = E.conv(E.View(x), pw, engine="f32", **kw).t
    launches = L.launch_count()
    yd = E.conv(E.View(x), pw, engine="tc", **kw).t
    assert L.launch_count() - launches == 2                     # cell GEMM + FIR pass
    monkeypatch.setenv("E4S_UPZ_DIRECT", "0")
    yh = E.conv(E.View(x), pw, engine="tc", **kw).t
    torch.cuda.synchronize()
    scale = float(y32.abs().max())
    assert float((yd - y32).abs().max()) < 2e-4 * max(scale, 1.0)
    assert float((yh - y32).abs().max()) < 2e-4 * max(scale, 1.0)

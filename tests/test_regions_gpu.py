"""GPU: masked layers through the halo kernel's region-job path (one job per (16x8 tile, region present)) must give the
same result as the per-row gather path, and both must match the oracle."""
import pytest
import torch

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("up", [False, True])
def test_region_jobs_match_gather_and_oracle(up):
    from e4s2024_b200 import engine as E
    from e4s2024_b200.stylegan2.model import StyledConv
    if not E.tc_available():
        pytest.skip("no tcgen05 engine")
    K, B, cin, cout, res = 5, 2, 128, 64, 32
    m = StyledConv(cin, cout, 3, 512, upsample=up, mask_op=True)
    sd = synth.synth_module_weights(m, seed=31)
    m = m.cuda()
    x = synth.randn("rj.x", (B, cin, res, res), 31)
    style = synth.randn("rj.s", (B, K, 512), 31)
    out_res = 2 * res if up else res
    lab = synth.blocky_labels(B, K, 128, cells=8, seed=31)           # 16-px cells at 128^2 -> few regions per tile
    mask = synth.onehot(lab, K)
    noise = synth.randn("rj.n", (1, 1, out_res, out_res), 31)
    ref = orc.styled_conv(x, style, mask, {k: v.cpu() for k, v in sd.items()}, "", upsample=up, mask_op=True, noise=noise)

    old = E.REGION_JOB_RATIO
    try:
        E.REGION_JOB_RATIO = 1e9                                       # force the region-job path
        key = (out_res, out_res, up)
        ctx = E.RegionCtx(mask.cuda(), [key])
        assert ctx.onehot and ctx.jobs_for(*key) is not None and ctx.region_jobs[key].count >= ctx.region_jobs[key].tiles
        from e4s2024_b200 import _lib as L
        from e4s2024_b200.stylegan2.model import StyleRows
        xin = E.View(L.nchw_to_nhwc(x.cuda()))
        y_rj = L.nhwc_to_nchw(m.run(xin, StyleRows.from_tensor(style.cuda()), ctx, noise.cuda()).t)
        E.REGION_JOB_RATIO = 0.0                                       # force the per-row gather path
        y_g = L.nhwc_to_nchw(m.run(xin, StyleRows.from_tensor(style.cuda()), ctx, noise.cuda()).t)
    finally:
        E.REGION_JOB_RATIO = old
    d_rj = float((y_rj.cpu() - ref).abs().max())
    d_g = float((y_g.cpu() - ref).abs().max())
    print(f"up={up}: region-job path vs oracle {d_rj:.3e}, gather path vs oracle {d_g:.3e}, jobs {ctx.region_jobs[key].count} / tiles {ctx.region_jobs[key].tiles}")
    assert d_rj < 5e-4 and d_g < 5e-4


def test_masked_mean_19_regions_vs_oracle():
    """get_per_comp_styleCode with more regions than one launch holds (the seg19 UI variant of the reference, run_UI_seg19.py):
    regions are processed 16 per launch; empty regions give zero vectors."""
    from e4s2024_b200 import _lib as L
    from oracle import e4s_oracle as orc
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(2, 40, 16, 16, generator=g)
    lab = torch.randint(0, 18, (2, 1, 64, 64), generator=g)            # region 18 stays empty
    mask = torch.zeros(2, 19, 64, 64).scatter_(1, lab, 1.0)
    codes = torch.full((2, 19, 48), 7.0, device="cuda")
    L.masked_mean(feat.permute(0, 2, 3, 1).contiguous().cuda(), 40, mask.cuda(), codes, 8)
    ref = orc.masked_region_mean(feat, mask)
    assert float((codes[:, :, 8:].cpu() - ref).abs().max()) < 1e-6
    assert float(codes[:, 18, 8:].abs().max()) == 0.0 and float((codes[:, :, :8] - 7.0).abs().max()) == 0.0


def test_masked_mean_bits_matches_plane_kernel_and_oracle():
    """The membership-bit-map pooling the encoder uses (one word per pixel, 8 pixel chunks) against the float-plane kernel and the
    oracle, with an OVERLAPPING soft mask (a pixel in two regions counts for both, like `mask != 0` in the reference)."""
    from e4s2024_b200 import _lib as L
    from oracle import e4s_oracle as orc
    g = torch.Generator().manual_seed(11)
    for (c, hw, k) in ((256, 64, 12), (40, 16, 19), (512, 32, 12)):
        feat = torch.randn(2, c, hw, hw, generator=g)
        lab = torch.randint(0, k - 1, (2, 1, 128, 128), generator=g)
        mask = torch.zeros(2, k, 128, 128).scatter_(1, lab, 1.0)
        mask[:, 1, 10:60, 20:90] = 0.3                                  # overlaps whatever is there
        fn = feat.permute(0, 2, 3, 1).contiguous().cuda()
        a = torch.zeros(2, k, c + 8, device="cuda")
        L.masked_mean_bits(fn, c, L.mask_member_bits(mask.cuda()), k, a, 8)
        ref = orc.masked_region_mean(feat, mask)
        assert float((a[:, :, 8:].cpu() - ref).abs().max()) < 1e-6
        b = torch.zeros(2, k, c + 8, device="cuda")
        L.masked_mean(fn, c, mask.cuda(), b, 8)
        assert float((a - b).abs().max()) < 1e-6
        solo = torch.zeros(1, k, c + 8, device="cuda")
        L.masked_mean_bits(fn[1:].contiguous(), c, L.mask_member_bits(mask[1:].cuda()), k, solo, 8)
        assert torch.equal(solo[0], a[1])                               # batch-invariant


def test_skinny_batched_linear_matches_fp64_and_is_batch_invariant():
    """The LocalMLP-shaped batched launch (few rows, large weights -> linear_skinny_batched_kernel) against torch fp64, for row counts
    on both sides of its 16-row pass, with the in_square / rsqrt form of the demodulation GEMM and a leaky-ReLU epilogue."""
    from e4s2024_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(12)
    w = torch.randn(1040, 1280, generator=g) / 36.0
    bias = torch.randn(1040, generator=g)
    pw = E.pack_linear_weight(w.cuda())
    assert pw.cin * pw.cout >= 512 * 1024
    outs = {}
    for rows in (1, 16, 37):
        x = torch.randn(40, 1280, generator=torch.Generator().manual_seed(13))[:rows].contiguous()
        y, p = E.linear_rows(x.cuda(), rows, 1280, 0, pw, bias=bias.cuda(), act=L.ACT_LRELU, slope=0.01, gain=1.0, launch=False)
        L.conv_batched([p])
        ref = torch.nn.functional.leaky_relu(x.double() @ w.double().t() + bias.double(), 0.01)
        assert float((y.cpu().double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        outs[rows] = y.cpu()
    assert torch.equal(outs[1][0], outs[16][0]) and torch.equal(outs[16][5], outs[37][5])


def test_fused_bicubic_is_bit_identical_to_the_two_pass_kernels():
    """e4s_bicubic_down_norm_f32: the single-kernel form (shared-memory tile, c_pad % 4 == 0) against the two-pass kernels it replaces
    (reached with c_pad = 5): same fmaf chains in the same order -> identical floats, on a size whose last tiles are partial."""
    from e4s2024_b200 import _lib as L
    from oracle import e4s_oracle as orc
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 176, 208, generator=g).cuda()
    for factor in (2, 4):
        taps = orc.bicubic_taps(factor).float().cuda()
        mean = torch.tensor([0.485, 0.456, 0.406]).cuda()
        std = torch.tensor([0.229, 0.224, 0.225]).cuda()
        fused = L.bicubic_down_norm(x, factor, taps, mean, std, 8)
        twopass = L.bicubic_down_norm(x, factor, taps, mean, std, 5)
        assert torch.equal(fused[..., :3], twopass[..., :3])
        assert float(fused[..., 3:].abs().max()) == 0.0 and float(twopass[..., 3:].abs().max()) == 0.0


def test_single_launch_chan_stats_matches_the_two_kernel_form_and_fp64():
    """e4s_chan_stats_f32: the one-launch form (c % 4 == 0; ticket counter, last block reduces) against the two-kernel form (c = 7 of the same
    buffer) and torch fp64; repeated calls reuse the zero-restored counters of the workspace."""
    from e4s2024_b200 import _lib as L
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(3, 40, 24, 8, generator=g) * 2 + 0.5).cuda()
    ref_m = x.double().mean((1, 2))
    ref_r = 1.0 / torch.sqrt(x.double().var((1, 2), unbiased=False) + 1e-5)
    for _ in range(3):
        m8, r8 = L.chan_stats(x, 8)
        m7, r7 = L.chan_stats(x, 7)
        assert float((m8.double() - ref_m).abs().max()) < 1e-6 and float((r8.double() / ref_r - 1).abs().max()) < 1e-6
        assert float((m8[:, :7] - m7).abs().max()) < 1e-6 and float((r8[:, :7] / r7 - 1).abs().max()) < 1e-6
    big = torch.randn(2, 256, 256, 64, generator=g).cuda()          # several slices per (sample, channel group)
    mb, rb = L.chan_stats(big, 64)
    assert float((mb.double() - big.double().mean((1, 2))).abs().max()) < 1e-6
    assert float((rb.double() * torch.sqrt(big.double().var((1, 2), unbiased=False) + 1e-5) - 1).abs().max()) < 1e-6

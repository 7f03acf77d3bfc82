"""Host-side launch helpers shared by the drop-in modules: weight packing, the conv launcher
(fills `struct E4SConv`), and the per-forward region context built from the mask.

Activations inside the engine are fp32 NHWC tensors [B,H,W,P] (P = pixel pitch >= channels);
the nn.Module boundary stays NCHW like the reference.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib as L

# engine selection for eligible convolutions: "tc" (tcgen05 bf16x3) or "f32" (CUDA-core fp32)
_ENGINE = os.environ.get("E4S_CONV_ENGINE", "tc")


PROFILE = None        # set to a list by bench.py to time every conv launch with CUDA events
PROFILE_STAGE = None  # label bench.py attaches to the launches it records ("parse" / "encoder" / "generator")
TC_UNBIAS_OVERRIDE = None   # tests/micro/acc_bias.py: E4SConv.tc_unbias for every tensor-core launch (< 0 = no correction)


def set_conv_engine(name: str):
    global _ENGINE
    assert name in ("tc", "f32")
    _ENGINE = name


def conv_engine() -> str:
    return _ENGINE


_PACK_ATTRS = ("_pack", "_e4s_pack", "_e4s_affine", "_wrgb", "_bias_cache")


def invalidate_packs(module: torch.nn.Module):
    """Drop every cached engine packing (packed / tensor-core weights, folded BatchNorm, ToRGB rows, MLP biases) under `module`.
    The caches are keyed on (data_ptr, _version) of the parameters, which in-place updates through `.data` do NOT change (the
    reference's Ranger optimiser and EMA `accumulate` write weights exactly that way): call this after such an update.
    `load_state_dict` on the drop-in modules calls it automatically."""
    for m in module.modules():
        for a in _PACK_ATTRS:
            if getattr(m, a, None) is not None:
                setattr(m, a, None)


def install_pack_invalidation(module: torch.nn.Module):
    """load_state_dict(...) -> invalidate_packs (registered by Generator / Net3 / FSEncoder_PSP / BiSeNet)."""
    module.register_load_state_dict_post_hook(lambda mod, incompatible: invalidate_packs(mod))


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


@dataclass
class PackedConv:
    """Weights in the e4s_conv_f32 layout [phases, K, cout_pad] (+ optional tensor-core image)."""
    w: torch.Tensor
    cin: int
    cout: int
    cout_pad: int
    kh: int
    kw: int
    phases: int = 1
    tc: Optional[torch.Tensor] = None
    tc9: Optional[torch.Tensor] = None   # up-convolutions: the 9 conv_transpose taps packed for e4s_conv_tc_upz
    fir: Optional[torch.Tensor] = None   # ... and the 4x4 blur kernel its finishing pass applies
    tc_fmt: int = L.TC_BF16          # operand format of the tensor-core image (E4SConv.tc_fmt)
    tc_out_scale: float = 1.0        # 1 / the power-of-two pre-scale of fp16 weights (E4SConv.tc_out_scale)

    @property
    def k(self) -> int:
        return self.kh * self.kw * self.cin


def halo_geometry_ok(cin: int, cout: int, hin: int, win: int, up2: bool, kh: int = 3, stride: int = 1) -> bool:
    """Mirror of tc_halo_geometry_ok (conv_tc_halo.cu): what the halo kernel takes."""
    if kh != 3 or stride != 1 or hin % 16 or win % 8:
        return False
    if not ((cin == 32 and cout == 32) or cin % 64 == 0):
        return False
    return (4 if up2 else 1) * min(cout, 256) <= 512 and tc_eligible(cin, cout)


def rgb_fusable(cin: int, cout: int, h: int, w: int) -> bool:
    """Can a same-resolution 3x3 layer carry the fused ToRGB tail (e4s_b200.h, struct E4SConv rgb*)?"""
    # cout <= 128: the halo kernel runs wider layers as 128-column n-tiles, and the tail needs a pixel's whole channel row in one job
    return _ENGINE == "tc" and tc_available() and cout <= 128 and halo_geometry_ok(cin, cout, h, w, False) and \
        os.environ.get("E4S_FUSE_RGB", "1") != "0"


def wide_eligible(cin: int, cout: int, hin: int, win: int, up2: bool) -> bool:
    """Mirror of tc_wide_eligible (conv_tc_wide.cu) for the StyledConv epilogue: geometry of the 512-column gather kernel."""
    if _ENGINE != "tc" or os.environ.get("E4S_TC_WIDE", "1") == "0" or cin % 64 or hin % 16 or win % 8:
        return False
    return cout % 128 == 0 if up2 else cout % 512 == 0


WIDE_OVER_REGION_JOBS = float(os.environ.get("E4S_WIDE_RATIO", "1.25"))


def splitk_factor(hout: int, wout: int, k: int) -> int:
    """K slices for the generator's smallest layers (e4s_conv_tc_splitk).  Chosen from the layer's geometry ONLY -- never from the batch --
    so that a sample computed alone and inside a batch sees the same summation order (bit-exact batch invariance)."""
    num_kc = (k + 63) // 64
    hw = hout * wout
    if os.environ.get("E4S_SPLITK", "1") == "0" or hw > 256 or num_kc < 16:
        return 1
    for ks in range(8 if hw <= 64 else 4, 1, -1):        # per-CTA fixed cost (~15 us) against 0.6-1 us per chunk: few, longer slices win
        if num_kc % ks == 0:
            return ks
    return 1


def upz_eligible(cin: int, cout: int, regional: bool = True) -> bool:
    """Mirror of e4s_conv_tc_upz's shape check: up-convolutions that can run as the conv_transpose cell GEMM + FIR pass --
    regional (masked) layers over (cell, region) rows, un-masked layers in the direct form (one row per cell, also cout 32 / 64)."""
    if _ENGINE != "tc" or os.environ.get("E4S_UPZ", "1") == "0" or cin % 64 or not tc_available():
        return False
    if regional:
        return cout % 128 == 0
    # Off by default: measured on B200 (tests/micro/upz_direct_time.py, profiles/r2_upz_direct.txt) the direct form LOSES on the un-masked
    # 512^2 / 1024^2 layers -- 2.5 vs 1.3 ms and 5.1 vs 2.0 ms against the poly-phase halo kernel.  With cin = 128 / 64 a 128-row tile has one
    # or two 64-channel groups of MMAs, so the one-tile-per-CTA kernel is all prologue / epilogue (10 us per tile), and the Z round trip
    # through HBM (2 x 2.1 GB at 1024^2) costs more than the 4x MMAs it saves.  It would take the halo kernel's persistent, overlapped
    # structure with the FIR applied in its epilogue to win here.
    return os.environ.get("E4S_UPZ_DIRECT", "0") == "1" and (cout in (32, 64) or cout % 128 == 0)


# run the (cell, region) form while it has at most this many rows per cell (its MMA work is rows x 9 against cells x 36 -- x72 on tiles
# with mixed phases -- for the poly-phase kernel); beyond that the poly-phase kernel runs (device-side predicate)
UPZ_RATIO = float(os.environ.get("E4S_UPZ_RATIO", "2.5"))


def tc_eligible(cin: int, cout: int) -> bool:
    return cin % 8 == 0 and (cout in (32, 64, 128) or (cout >= 256 and cout % 256 == 0))


def tc_available() -> bool:
    """True when the library was built with the tcgen05 engine."""
    return int(L.lib().e4s_pack_weights_tc_bytes(1, 64, 32)) > 0


def _finish_pack(w_pkc: torch.Tensor, cin: int, cout: int, kh: int, kw: int, phases: int, want_tc: bool,
                 tc_fmt: int = L.TC_BF16) -> PackedConv:
    """w_pkc: [phases, K, cout_pad] fp32 from L.pack_conv_weights / L.pack_upconv_weights."""
    cout_pad = w_pkc.shape[2]
    pc = PackedConv(w_pkc, cin, cout, cout_pad, kh, kw, phases)
    if want_tc and w_pkc.is_cuda and tc_eligible(cin, cout) and tc_available():
        scale = 1.0
        if tc_fmt == L.TC_F16:
            # fp16 operands: bring max|w| to [2^13, 2^14) with a power of two (exact), so that the hi parts of all but the
            # tiniest weights stay normal fp16 numbers; the launch undoes it through E4SConv.tc_out_scale
            mx = float(w_pkc.abs().max())
            if mx > 0.0 and math.isfinite(mx):
                scale = 2.0 ** (13 - math.floor(math.log2(mx)))
        pc.tc = L.pack_weights_tc(w_pkc, phases, pc.k, cin, cout, cout_pad, tc_fmt, scale)
        pc.tc_fmt, pc.tc_out_scale = tc_fmt, 1.0 / scale
    return pc


def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None, want_tc: bool = True, tc_fmt: int = L.TC_BF16,
                     scale: float = 1.0, sumsq: bool = False) -> PackedConv:
    """w [Co,Ci,kh,kw] -> k = (ky*kw + kx)*Ci_pad + ci rows, co contiguous (one kernel: e4s_pack_conv_weights_f32).
    sumsq: the [Ci] x [Co] matrix sum_taps (scale*w)^2 instead (demodulation table GEMM)."""
    co, ci, kh, kw = w.shape
    cin = ci if cin_pad is None else cin_pad
    m = L.pack_conv_weights(w.detach().contiguous().float(), cin, pad_to(co, 4), scale, sumsq)
    if sumsq:
        return _finish_pack(m, cin, co, 1, 1, 1, False)
    return _finish_pack(m, cin, co, kh, kw, 1, want_tc, tc_fmt)


def pack_linear_weight(w: torch.Tensor, want_tc: bool = False, scale: float = 1.0) -> PackedConv:
    """w [out,in] (F.linear) as a 1x1 conv."""
    return pack_conv_weight(w.detach()[:, :, None, None], want_tc=want_tc, scale=scale)


def pack_up_weight(w: torch.Tensor, fir: torch.Tensor, want_tc: bool = True, scale: float = 1.0) -> PackedConv:
    """conv_transpose2d(stride 2, weight w[Co,Ci,3,3] used as [Ci,Co,3,3]) followed by upfirdn2d(fir 4x4, pad=(1,1)) as four 3x3
    phase filters applied at input resolution (one kernel: e4s_pack_upconv_weights_f32)."""
    co, ci = w.shape[:2]
    wf, ff = w.detach().contiguous().float(), fir.detach().contiguous().float().to(w.device)
    m = L.pack_upconv_weights(wf, ff, pad_to(co, 4), scale)
    pc = _finish_pack(m, ci, co, 3, 3, 4, want_tc)
    if pc.tc is not None and (upz_eligible(ci, co, True) or upz_eligible(ci, co, False)):
        m9 = L.pack_convt_weights(wf, pad_to(co, 4), scale)
        pc.tc9 = L.pack_weights_tc(m9, 9, ci, ci, co, pad_to(co, 4))
        pc.fir = ff
    return pc


class View:
    """An NHWC activation: tensor [B,H,W,P] + channel window [coff, coff+c)."""
    __slots__ = ("t", "c", "coff")

    def __init__(self, t: torch.Tensor, c: Optional[int] = None, coff: int = 0):
        assert t.dim() == 4 and t.is_contiguous() and t.dtype == torch.float32
        self.t, self.coff = t, coff
        self.c = t.shape[3] - coff if c is None else c

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + 4 * self.coff

    @property
    def pitch(self) -> int:
        return self.t.shape[3]

    @property
    def bhw(self) -> Tuple[int, int, int]:
        return self.t.shape[0], self.t.shape[1], self.t.shape[2]


def new_nhwc(b, h, w, c, device) -> torch.Tensor:
    return torch.empty(b, h, w, c, device=device, dtype=torch.float32)


def conv(x: View, pw: PackedConv, *, stride=1, pad=None, up2=False, in_shift=0, in_stats=None, in_square=False,
         smod=None, demod=None, labels=None, regions=1, smod_off=0, demod_off=0, pixw=None, ch_scale=None, ch_shift=None,
         noise=None, noise_w=None, res: Optional[View] = None, res_after_act=False, act=L.ACT_NONE, slope=0.0, gain=1.0,
         prelu=None, out: Optional[View] = None, accumulate=False, engine: Optional[str] = None,
         region_jobs: Optional["RegionJobs"] = None, rgb: Optional[dict] = None, store_out: bool = True,
         upz: Optional["UpzRows"] = None) -> Optional[View]:
    """Launch one fused convolution (see struct E4SConv).  Returns the output view.
    `rgb` = {"rgb": [B,3,H,W], "w": [3,cout], "smod": [B,cout], "bias": [3]|None, "skip": [B,3,H/2,W/2]|None, "fir": [4,4]|None}
    adds the fused ToRGB tail (halo kernel only, see rgb_fusable); with store_out=False the activations are never written."""
    b, hin, win = x.bhw
    assert x.c == pw.cin, (x.c, pw.cin)
    if x.t.is_cuda:
        L.check_device(x.t)
    pad = (pw.kh // 2) if pad is None else pad
    if up2:
        hout, wout = 2 * hin, 2 * win
    else:
        hv, wv = hin << in_shift, win << in_shift
        hout = (hv + 2 * pad - pw.kh) // stride + 1
        wout = (wv + 2 * pad - pw.kw) // stride + 1
    if not store_out:
        assert rgb is not None and out is None
    elif out is None:
        out = View(new_nhwc(b, hout, wout, pw.cout, x.t.device))
    if out is not None:
        assert out.bhw == (b, hout, wout) and out.c == pw.cout, (out.bhw, (b, hout, wout), out.c, pw.cout)
    p = L.E4SConv()
    p.x, p.x_pitch = x.ptr, x.pitch
    p.batch, p.hin, p.win, p.cin = b, hin, win, pw.cin
    p.in_shift, p.in_square = in_shift, int(in_square)
    p.w = pw.w.data_ptr()
    p.cout, p.cout_pad = pw.cout, pw.cout_pad
    p.kh, p.kw, p.stride, p.pad = pw.kh, pw.kw, stride, pad
    p.mode = L.CONV_UP2 if up2 else L.CONV_NORMAL
    p.hout, p.wout = hout, wout
    if in_stats is not None:
        p.in_mean, p.in_rstd = in_stats[0].data_ptr(), in_stats[1].data_ptr()
    p.regions = regions
    if smod is not None:
        p.smod = smod.data_ptr() + 4 * smod_off
    if demod is not None:
        p.demod = demod.data_ptr() + 4 * demod_off
    if labels is not None:
        p.labels = labels.data_ptr()
        p.lab_h, p.lab_w = labels.shape[1], labels.shape[2]
    if pixw is not None:                                   # (tensor [B,K,Hm,Wm], region index)
        m, ridx = pixw
        p.pixw = m.data_ptr() + 4 * ridx * m.shape[2] * m.shape[3]
        p.pixw_sb = m.shape[1] * m.shape[2] * m.shape[3]
        p.lab_h, p.lab_w = m.shape[2], m.shape[3]
    if ch_scale is not None:
        p.ch_scale = ch_scale.data_ptr()
    if ch_shift is not None:
        p.ch_shift = ch_shift.data_ptr()
    if noise is not None:
        nb, nc, nh, nw_ = noise.shape
        assert (nh, nw_) == (hout, wout) and nb in (1, b) and nc in (1, pw.cout) and noise.is_contiguous(), noise.shape
        p.noise, p.noise_w = noise.data_ptr(), noise_w.data_ptr()
        p.noise_sb = 0 if nb == 1 else nc * nh * nw_
        p.noise_sc = 0 if nc == 1 else nh * nw_
    if res is not None:
        assert res.bhw == (b, hout, wout)
        p.res, p.res_pitch, p.res_after_act = res.ptr, res.pitch, int(res_after_act)
    p.act, p.act_slope, p.act_gain = act, slope, gain
    if prelu is not None:
        p.act_prelu = prelu.data_ptr()
    if out is not None:
        p.out, p.out_pitch, p.accumulate = out.ptr, out.pitch, int(accumulate)
    else:
        p.out_pitch = pw.cout
    if rgb is not None:
        assert rgb["rgb"].shape == (b, 3, hout, wout) and rgb["rgb"].is_contiguous() and rgb["w"].shape == (3, pw.cout)
        assert rgb["smod"].numel() == b * pw.cout and rgb["smod"].is_contiguous() and rgb["w"].is_contiguous()
        p.rgb, p.rgb_w, p.rgb_smod = rgb["rgb"].data_ptr(), rgb["w"].data_ptr(), rgb["smod"].data_ptr()
        if rgb.get("bias") is not None:
            p.rgb_bias = rgb["bias"].data_ptr()
        if rgb.get("skip") is not None:
            assert rgb["skip"].shape == (b, 3, hout // 2, wout // 2) and rgb["skip"].is_contiguous()
            p.rgb_skip, p.rgb_fir = rgb["skip"].data_ptr(), rgb["fir"].data_ptr()
    eng = engine or _ENGINE
    use_tc = eng == "tc" and pw.tc is not None
    use_rj = use_tc and region_jobs is not None and labels is not None
    upz_epi_ok = (noise is None or noise.shape[1] == 1) and ch_scale is None and res is None and pixw is None and not accumulate and \
        act in (L.ACT_NONE, L.ACT_LRELU) and in_stats is None and rgb is None and pw.tc9 is not None and up2 and use_tc and in_shift == 0
    use_upz = upz_epi_ok and upz is not None and labels is not None and upz_eligible(pw.cin, pw.cout, True)
    # un-masked up-convolution (one style per sample): the same cell GEMM without a row list
    use_upz_direct = upz_epi_ok and labels is None and regions == 1 and demod_off == 0 and smod_off == 0 and upz_eligible(pw.cin, pw.cout, False)
    if use_tc:
        p.tc_fmt, p.tc_out_scale = pw.tc_fmt, pw.tc_out_scale
        if TC_UNBIAS_OVERRIDE is not None:
            p.tc_unbias = TC_UNBIAS_OVERRIDE

    def launch_upz():
        z = torch.empty(upz.max_rows * 4 * pw.cout, device=x.t.device, dtype=torch.float32)
        L.conv_upz(p, pw.tc9, pw.fir, upz.cells, upz.rows, upz.count_dev, upz.max_rows, z)

    # the generator's 4^2 - 16^2 modulated layers: a handful of output tiles, 72 K chunks each -> split-K (deterministic two-pass reduction)
    ksplit = 1
    if use_tc and not up2 and smod is not None and pw.tc_fmt == L.TC_BF16 and stride == 1 and in_shift == 0 and in_stats is None and rgb is None and \
            res is None and pixw is None and not accumulate and (noise is None or noise.shape[1] == 1) and prelu is None and \
            act in (L.ACT_NONE, L.ACT_LRELU, L.ACT_RELU) and not in_square and out is not None:
        ksplit = splitk_factor(hout, wout, pw.k)

    def launch():
        if ksplit > 1:
            L.conv_splitk(p, pw.tc, ksplit)
        elif use_upz_direct:
            n_cells = b * (hin + 1) * (win + 1)
            z = torch.empty(n_cells * 4 * pw.cout, device=x.t.device, dtype=torch.float32)
            L.conv_upz(p, pw.tc9, pw.fir, None, None, None, n_cells, z)
        elif use_upz and upz.count is None:
            # lazy region context: the row count on the device runs either the (cell, region) form or the poly-phase kernel
            p.pred_count, p.pred_limit = upz.count_dev.data_ptr(), upz.limit
            p.pred_run_if_gt = 0
            launch_upz()
            p.pred_run_if_gt = 1
            L.conv(p, pw.tc)
        elif use_upz and upz.count <= upz.limit:
            launch_upz()
        elif use_rj and region_jobs.count is None:
            # lazy region context: enqueue both candidates, the job count on the device runs exactly one of them
            p.pred_count, p.pred_limit = region_jobs.count_dev.data_ptr(), region_jobs.limit
            p.pred_run_if_gt = 0
            L.conv_regions(p, pw.tc, region_jobs.jobs, region_jobs.count_dev, region_jobs.jobs.shape[0])
            p.pred_run_if_gt = 1
            L.conv(p, pw.tc)
        elif use_rj:
            L.conv_regions(p, pw.tc, region_jobs.jobs, region_jobs.count_dev, region_jobs.count)
        else:
            L.conv(p, pw.tc if use_tc else None)

    if PROFILE is None:
        launch()
        return out
    # bench.py's per-launch timing pass: CUDA events on the launching stream around this one kernel
    m_exec = b * hout * wout
    alg = 2.0 * (m_exec / 4 if up2 else m_exec) * pw.k * pw.cout      # conv_transpose counted at input resolution
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launch()
    ev1.record()
    PROFILE.append({"stage": PROFILE_STAGE, "engine": "tc" if use_tc else "f32", "fmt": ("f16" if pw.tc_fmt == L.TC_F16 else "bf16") if use_tc else "f32",
                    "cin": pw.cin, "hout": hout, "stride": stride, "kh": pw.kh, "region_jobs": (region_jobs.count if region_jobs.count is not None else -1) if use_rj else 0, "alg_flops": alg, "exec_flops": 2.0 * m_exec * pw.k * pw.cout,
                    "m": m_exec, "k": pw.k, "n": pw.cout, "up2": bool(up2), "ev": (ev0, ev1), "rgb": rgb is not None,
                    "upz_rows_per_cell": (float(upz.count_dev.item()) / upz.cells_total) if use_upz else 0.0,
                    "bytes": 4.0 * (b * hin * win * pw.cin + (m_exec * pw.cout if out is not None else 0) + (3 * m_exec if rgb is not None else 0))})
    return out


def linear_rows(x_ptr_tensor: torch.Tensor, rows: int, row_stride: int, offset: int, pw: PackedConv, bias=None,
                act=L.ACT_NONE, slope=0.0, gain=1.0, in_square=False, out: Optional[torch.Tensor] = None,
                engine: Optional[str] = None, launch: bool = True):
    """y[r,:] = act(W x[r,:] + bias) for `rows` vectors of length pw.cin starting at element `offset`
    of `x_ptr_tensor`, consecutive rows `row_stride` elements apart (all in floats)."""
    dev = x_ptr_tensor.device
    if out is None:
        out = torch.empty(rows, pw.cout, device=dev, dtype=torch.float32)
    p = L.E4SConv()
    p.x, p.x_pitch = x_ptr_tensor.data_ptr() + 4 * offset, row_stride
    p.batch, p.hin, p.win, p.cin = rows, 1, 1, pw.cin
    p.in_square = int(in_square)
    p.w = pw.w.data_ptr()
    p.cout, p.cout_pad = pw.cout, pw.cout_pad
    p.kh = p.kw = p.stride = 1
    p.pad, p.mode = 0, L.CONV_NORMAL
    p.hout = p.wout = 1
    p.regions = 1
    if bias is not None:
        p.ch_shift = bias.data_ptr()
    p.act, p.act_slope, p.act_gain = act, slope, gain
    p.out, p.out_pitch = out.data_ptr(), out.stride(0)
    if not launch:                      # caller batches several problems into one launch (L.conv_batched)
        return out, p
    eng = engine or "f32"
    L.conv(p, pw.tc if (eng == "tc" and pw.tc is not None) else None)
    return out


REGION_JOB_RATIO = float(os.environ.get("E4S_REGION_RATIO", "3.0"))


@dataclass
class RegionJobs:
    """(16x8 tile, region present) job list of one masked layer resolution (e4s_region_tile_jobs)."""
    jobs: torch.Tensor
    count_dev: torch.Tensor
    count: Optional[int]          # None: not read back (lazy context) -> the choice is made on the device (E4SConv.pred_*)
    tiles: int
    limit: int = 0                # lazy: run the region-job kernel iff count <= limit, the per-row kernel otherwise


@dataclass
class UpzRows:
    """(cell, region) row list of one masked up-convolution resolution (e4s_upz_build_rows)."""
    cells: torch.Tensor
    rows: torch.Tensor
    count_dev: torch.Tensor
    count: Optional[int]          # None: not read back (lazy context) -> device-side predicate
    max_rows: int
    limit: int
    cells_total: int


class RegionCtx:
    """Per-forward view of the mask [B,K,Hm,Wm]: u8 label map when every pixel has at most one
    region with weight exactly 1 (the pipelines' one-hot masks), else the generic float path.
    `job_keys` = [(hout, wout, up2)] of the masked 3x3 layers whose geometry the halo kernel takes: their
    per-tile region job lists are built here and the counts come back in the same (single) D2H read as the flag."""

    def __init__(self, mask: torch.Tensor, job_keys=(), lazy: bool = False, host_flag: Optional[torch.Tensor] = None, upz_keys=()):
        """lazy=True (Generator.forward): nothing is read back while the forward is being enqueued.  The mask is ASSUMED
        one-hot (checked by `verify()` after the last launch; the caller re-runs on the generic path if not) and every
        region-job decision is a device-side launch predicate."""
        if mask.dim() != 4:
            raise L.E4SError("mask must be [B,K,H,W]")
        self.mask = mask.contiguous().float()
        self.k = mask.shape[1]
        keys = list(dict.fromkeys(job_keys)) if (self.mask.is_cuda and self.k <= 32 and tc_available()) else []
        ukeys = list(dict.fromkeys(upz_keys)) if (self.mask.is_cuda and self.k <= 32 and tc_available()) else []
        meta = torch.zeros(1 + len(keys) + len(ukeys), device=self.mask.device, dtype=torch.int32)
        self.labels, _ = L.mask_labels(self.mask, meta[0:1])
        lists = [L.region_tile_jobs(self.labels, h, w, up, self.k, meta[1 + i:2 + i]) for i, (h, w, up) in enumerate(keys)]
        ulists = []
        for i, (hin, win) in enumerate(ukeys):
            cells_total = self.mask.shape[0] * (hin + 1) * (win + 1)
            limit = int(UPZ_RATIO * cells_total)
            max_rows = pad_to(limit, 128)
            slot = meta[1 + len(keys) + i:2 + len(keys) + i]
            cells, rows = L.upz_build_rows(self.labels, hin, win, max_rows, slot)
            ulists.append((cells, rows, slot, max_rows, limit, cells_total))
        self.lazy = bool(lazy) and self.mask.is_cuda
        if self.lazy:
            # host_flag: a pre-allocated pinned int32 buffer (CUDA-graph capture: nothing may be allocated on the host or waited
            # for inside the captured region; the caller reads host_flag[0] after the replay has completed)
            self._host = host_flag if host_flag is not None else torch.empty(meta.shape, dtype=torch.int32, pin_memory=True)
            self._host[:meta.numel()].copy_(meta, non_blocking=True)
            self._ev = None
            if host_flag is None:
                self._ev = torch.cuda.Event()
                self._ev.record()
            host = None
            self.onehot = True                          # speculation; see verify()
        else:
            host = meta.cpu()                           # one small D2H read per forward
            self.onehot = int(host[0]) == 0
        self.regions = self.k
        self.region_jobs = {}
        b = self.mask.shape[0]
        for i, (key, jl) in enumerate(zip(keys, lists)):
            h, w, up = key
            gh, gw = (h // 2, w // 2) if up else (h, w)
            self.region_jobs[key] = RegionJobs(jl, meta[1 + i:2 + i], None if host is None else int(host[1 + i]),
                                               b * (gh // 16) * (gw // 8))
        self.upz = {}
        for i, (key, (cells, rows, slot, max_rows, limit, cells_total)) in enumerate(zip(ukeys, ulists)):
            self.upz[key] = UpzRows(cells, rows, slot, None if host is None else int(host[1 + len(keys) + i]), max_rows, limit, cells_total)

    def verify(self) -> bool:
        """lazy contexts: was the one-hot assumption right?  (waits only for the tiny copy issued before the first layer)"""
        if not self.lazy or self._ev is None:           # graph capture: the owner of host_flag checks it after the replay
            return True
        self._ev.synchronize()
        return int(self._host[0]) == 0

    def jobs_for(self, hout: int, wout: int, up2: bool, wide_ok: bool = False) -> Optional["RegionJobs"]:
        """The job list if this resolution has few enough regions per tile for the halo kernel to win
        (`wide_ok`: the layer can run on the 512-column gather kernel, which costs one pass whatever the mask)."""
        rj = self.region_jobs.get((hout, wout, bool(up2)))
        if rj is None or not self.onehot:
            return None
        ratio = min(REGION_JOB_RATIO, WIDE_OVER_REGION_JOBS) if wide_ok else REGION_JOB_RATIO
        if rj.count is None:                            # lazy: both kernels are enqueued, the device count picks one
            rj.limit = int(ratio * rj.tiles)
            return rj
        if rj.count <= 0 or rj.count > ratio * rj.tiles:
            return None
        return rj

    def upz_for(self, hin: int, win: int) -> Optional["UpzRows"]:
        """The (cell, region) row list of an up-convolution with hin x win input, if one was requested for this forward."""
        return self.upz.get((hin, win)) if self.onehot else None

    @property
    def lab_hw(self):
        return self.labels.shape[1], self.labels.shape[2]

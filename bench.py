#!/usr/bin/env python
"""bench.py -- 1024^2 faces/sec of the E4S swap hot path on N B200s (BASELINE.json metric).

Workload (config.workload): the per-GPU shard of BASELINE.json configs[4] -- 16 faces per GPU through the FULL path
    uint8 1024^2 image -> TO_TENSOR/NORMALIZE -> bicubic 1024->512 + BiSeNet + argmax + seg19->12 LUT -> one-hot
    -> Net3 (regional style encoder, 12 LocalMLPs, mask-guided StyleGAN2 1024^2, K=12, rl=13) -> tensor2im (uint8)
    [-> NCCL all-gather of the uint8 images + label maps when N > 1]
through `e4s2024_b200.sharding.SwapHotPath` (the public entry), synthetic weights and FFHQ-shaped synthetic images.
A "step" = one pass of that path over the 16-face shard of every rank.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this framework (CUDA, libe4s_b200.so)
  python bench.py --impl reference ...                         # the reference algorithm on the host CPU cores

value    : whole-job faces/s, inputs (uint8 images) resident in HBM, CUDA events, max over ranks, gathers inside the timed region.
e2e      : same metric with HOST buffers: pinned uint8 images in, uint8 images + u8 label maps out, every step, through
           serving.HostPipeline (double-buffered copies), timed from the first H2D to the last D2H.
roofline : the convolution launches of one step (the tcgen05 kernels that carry 405.5 GFLOP/face), per stage.
Extra keys (N = 1): configs[1] generator-only (three mask families), configs[2] (B=32 Net3), configs[3] (B=64 BiSeNet),
           gpu_reference (the reference algorithm as torch/cuDNN ops on the same B200, TF32 off / on), parity, cpu_baseline.
See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "1024^2 faces/sec (StyleGAN2 synthesis+encoder+parsing: full swap hot path)"
BATCH, SIZE, K, RL, SPLIT = 16, 1024, 12, 13, 5
# algorithmic GFLOP per face (each output pixel once, conv_transpose at input resolution)
ALG_GFLOP = {"parse": 27.54, "encoder": 229.34, "mlps": 0.10, "generator": 148.52}
ALG_GFLOP_PER_FACE = 405.5
LABEL_TIE_RULE = "labels may differ from the CPU oracle only where the oracle's own top-2 logit margin is < 2e-5 x max|logit| (fp32 reassociation noise)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "hbm_gbs": d["hbm_gbs"],
                "source": "MEASURED_PEAKS.json (sustained bf16 cuBLAS rate: the kernels are timed inside a long step; copy bandwidth)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def net3_opts():
    # the fields Net3 reads from the reference's option object (options/our_swap_face_pipeline_options.py:12-18,50)
    return types.SimpleNamespace(fsencoder_type="psp", remaining_layer_idx=RL, num_seg_cls=K, out_size=SIZE, train_G=False,
                                 start_from_latent_avg=True, learn_in_w=False)


def bench_config(world, engine="tc"):
    """`config` of the JSON line: the same dict for both arms (the reference arm adds what its bounded sample was)."""
    return {"workload": "configs[4] per-GPU shard (the configuration the metric is quoted on): 16 faces/GPU, full swap hot path "
                        "(bicubic 1024->512 + BiSeNet + argmax/LUT -> one-hot -> Net3: encoder, 12 LocalMLPs, StyleGAN2 1024^2 K=12 rl=13)",
            "batch_per_gpu": BATCH, "global_batch": BATCH * world, "size": SIZE, "regions": K, "remaining_layer_idx": RL,
            "inputs": "uint8 HWC 1024^2 images (low-pass noise, FFHQ-shaped); masks come from the parser on those images",
            "noise": "registered buffers (randomize_noise=False)",
            "parallelism": f"batch-sharded x{world}" + (" + NCCL all_gather of the uint8 images and label maps (async, double-buffered)" if world > 1 else ""),
            "conv_engine": engine, "label_tie_rule": LABEL_TIE_RULE,
            "l2": "working set (>2 GB of activations per step) exceeds the 126 MB L2; no flush needed"}


def build_path(dev, seed_net=9, seed_seg=10):
    from e4s2024_b200 import synth
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    from e4s2024_b200.networks import Net3
    from e4s2024_b200.sharding import SwapHotPath
    net = Net3(net3_opts())
    sd_net = synth.synth_module_weights(net, seed=seed_net)
    net = net.to(dev).eval()
    la = synth.randn("net3.latent_avg", (18, 512), seed_net, 0.1)
    net.latent_avg = la.to(dev)
    parser = FaceParser(seg_ckpt=None, size=SIZE, device=str(dev))
    sd_seg = synth.synth_module_weights(parser.seg, seed=seed_seg)
    parser.seg.to(dev)
    return SwapHotPath(net, parser, K), {"net": sd_net, "seg": sd_seg, "latent_avg": la}


def thread_candidates():
    """torch/oneDNN does not scale to every hardware thread on big hosts (128 threads were 6x slower than 8 on the
    first box), so the CPU arm tries a few thread counts and keeps the fastest; `cores` reports the one used."""
    n = os.cpu_count() or 1
    return sorted({n, min(n, 64), min(n, 32), min(n, 16)}, reverse=True)


def oracle_chain(sds, img_u8, labels_override=None):
    """The reference algorithm for one batch on whatever device the tensors live on (oracle/e4s_oracle.py): TO_TENSOR / NORMALIZE,
    face_parse (bicubic + BiSeNet + argmax + LUT), labelMap2OneHot, Net3.forward (encoder, MLPs, K grouped convs per masked layer),
    tensor2im.  -> (image fp32 [B,3,S,S], labels u8 [B,512,512] numpy, logits-free)."""
    from oracle import e4s_oracle as orc
    x01 = img_u8.permute(0, 3, 1, 2).float().div(255)
    x = (x01 - 0.5) / 0.5
    labels = orc.face_parse(sds["seg"], x01)
    lab_t = torch.from_numpy(labels if labels_override is None else labels_override).to(x.device).long()[:, None]
    mask = orc.label_to_onehot(lab_t, K)
    image = orc.net3_forward(sds["net"], x, mask, sds["latent_avg"], out_size=SIZE, remaining_layer_idx=RL)[0]
    return image, labels


def cpu_state(sds):
    return {"net": {k: v.cpu() for k, v in sds["net"].items()}, "seg": {k: v.cpu() for k, v in sds["seg"].items()},
            "latent_avg": sds["latent_avg"].cpu()}


def cpu_baseline(sds, img_u8_1, gpu_labels_1=None):
    """The oracle (CPU restatement of the reference algorithm) for ONE face of the workload on the host cores; also returns its
    outputs so the same run reports parity of the GPU path against it."""
    sd = cpu_state(sds)
    best, best_t, tried, out = None, None, {}, None
    with torch.no_grad():
        for t in thread_candidates():
            torch.set_num_threads(t)
            t0 = time.perf_counter()
            out = oracle_chain(sd, img_u8_1)
            dt = time.perf_counter() - t0
            tried[t] = round(dt, 2)
            if best is None or dt < best:
                best, best_t = dt, t
        image, labels = out
        if gpu_labels_1 is not None and (labels != gpu_labels_1).any():
            # image parity is judged on the same mask: re-run the oracle's Net3 on the GPU path's label map (the label map itself
            # is compared separately, under the tie rule)
            torch.set_num_threads(best_t)
            image, _ = oracle_chain(sd, img_u8_1, labels_override=gpu_labels_1)
    return ({"value": 1.0 / best, "unit": "faces/s", "cores": best_t, "kind": "port",
             "sample": f"1 face of the workload through the full path (parse + one-hot + Net3), oracle/e4s_oracle.py fp32; seconds per thread count {tried}"},
            image, labels)


def label_parity(sds, img_u8_1, gpu_labels_1):
    """Flip count and the oracle's top-2 margin at the flips for one face."""
    from oracle import e4s_oracle as orc
    sd = {k: v.cpu() for k, v in sds["seg"].items()}
    with torch.no_grad():
        x01 = img_u8_1.permute(0, 3, 1, 2).float().div(255)
        logits = orc.bisenet_forward(sd, orc.parser_preprocess(x01, SIZE))[0]
    lab19 = logits.argmax(1)
    ref12 = torch.from_numpy(orc.SEG19_TO_SEG12)[lab19.long()]
    bad = ref12 != torch.from_numpy(gpu_labels_1).to(ref12.dtype)
    top2 = torch.topk(logits, 2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    scale = float(logits.abs().max())
    return {"labels_compared": int(bad.numel()), "label_flips_vs_cpu_oracle": int(bad.sum()),
            "max_oracle_margin_at_a_flip": float(margin[bad].max()) if bad.any() else 0.0,
            "tie_threshold": 2e-5 * scale, "logit_absmax": scale, "tie_rule": LABEL_TIE_RULE}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from e4s2024_b200 import synth
    from e4s2024_b200.face_parsing.model import bisenet_state_shapes
    from e4s2024_b200.networks import net3_state_shapes
    cfg = bench_config(max(int(os.environ.get("WORLD_SIZE", "1")), 1), engine="cpu")
    cfg["sample"] = ("1 face per step (bounded sample of the 16-face shard) through the full path, CPU reference algorithm; the warm-up "
                     "steps double as a search over host thread counts (oneDNN does not scale to every hardware thread)")
    sd = {"net": synth.fill_state_dict(net3_state_shapes(net3_opts()), seed=9), "seg": synth.fill_state_dict(bisenet_state_shapes(19), seed=10),
          "latent_avg": synth.randn("net3.latent_avg", (18, 512), 9, 0.1)}
    img = synth.smooth_image_u8("swap.img", 1, SIZE, 13)
    step = lambda: oracle_chain(sd, img)
    cands = thread_candidates()
    threads, best = cands[0], None
    with torch.no_grad():
        for i in range(max(args.warmup, 1)):          # warm-up steps double as the thread-count search
            t = cands[i % len(cands)]
            torch.set_num_threads(t)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, threads = dt, t
        torch.set_num_threads(threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    v = args.steps / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "faces/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": cfg, "gpu_launches": 0,
                      "cpu_baseline": {"value": v, "unit": "faces/s", "cores": threads, "kind": "port",
                                       "sample": "1 face per step, oracle/e4s_oracle.py: face_parse + one-hot + net3_forward"},
                      "e2e": {"value": v, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def time_steps(fn, steps, warmup=2):
    """ms per call of fn (CUDA events on the current stream, synchronised on both sides)."""
    r = None
    for _ in range(warmup):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, r


def stage_breakdown(hot, img_u8_d, steps=5):
    """The stages of one step, each timed alone (CUDA events), plus the per-launch pass for the roofline."""
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    net, parser = hot.net, hot.parser
    ms = {}
    ms["im2tensor"], (img01, img) = time_steps(lambda: L.im2tensor(img_u8_d), steps)
    ms["parse"], lab = time_steps(lambda: parser.parse_batch(img01), steps)
    ms["onehot"], mask = time_steps(lambda: L.labels_to_onehot(lab, K), steps)
    ms["encoder"], (vec, _) = time_steps(lambda: net.get_style_vectors(img, mask), steps)
    ms["mlps"], codes = time_steps(lambda: net.cal_style_codes(vec), steps)
    ms["generator"], (out, _, _) = time_steps(lambda: net.gen_img(None, codes, mask, randomize_noise=False), steps)
    ms["tensor2im"], _ = time_steps(lambda: L.tensor2im_u8(out, True), steps)
    # per-launch pass: CUDA events around every convolution launch (engine.conv), tagged with its stage
    E.PROFILE = []
    for stage, fn in (("parse", lambda: parser.parse_batch(img01)), ("encoder", lambda: net.get_style_vectors(img, mask)),
                      ("generator", lambda: net.gen_img(None, codes, mask, randomize_noise=False))):
        E.PROFILE_STAGE = stage
        fn()
    torch.cuda.synchronize()
    prof, E.PROFILE, E.PROFILE_STAGE = E.PROFILE, None, None
    # how the parser's masks tile: region jobs per 16x8 tile (masked same-resolution layers) and (cell, region) rows per cell (masked up-convolutions)
    ctx = E.RegionCtx(mask, net.G._region_job_keys(), lazy=False, upz_keys=net.G._upz_keys())
    ms["mask_region_jobs_per_tile"] = {f"{h}x{w}{'_up' if up else ''}": round(rj.count / max(rj.tiles, 1), 3) for (h, w, up), rj in ctx.region_jobs.items()}
    ms["mask_upconv_rows_per_cell"] = {f"{h}x{w}->{2 * h}x{2 * w}": round(uz.count / max(uz.cells_total, 1), 3) for (h, w), uz in ctx.upz.items()}
    return ms, prof


def roofline_from(prof, batch, step_ms, traffic):
    pk = peaks()
    by = {}
    for r in prof:
        d = by.setdefault(r["stage"], {"ms": 0.0, "alg": 0.0, "exec": 0.0, "n": 0})
        d["ms"] += r["ev"][0].elapsed_time(r["ev"][1])
        d["alg"] += r["alg_flops"]
        d["exec"] += r["exec_flops"]
        d["n"] += 1
    conv_ms = sum(d["ms"] for d in by.values())
    if conv_ms <= 0:
        return None
    alg = ALG_GFLOP_PER_FACE * 1e9 * batch                     # SURVEY 8(d) per-face figure x faces per step
    counted = sum(d["alg"] for d in by.values())
    ach = alg / (conv_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_tc_{halo,wide,gather} (tcgen05, 3-pass hi/lo split: bf16 in the encoder / generator, fp16 in BiSeNet): "
                                         "every convolution launch of one step",
            "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": traffic,
            "peak_source": pk["source"], "launches": sum(d["n"] for d in by.values()), "kernel_ms_per_step": conv_ms,
            "conv_share_of_step": conv_ms / step_ms if step_ms else None,
            "alg_gflop_per_face": ALG_GFLOP_PER_FACE, "alg_gflop_per_face_counted_from_launches": counted / batch / 1e9,
            "executed_tflops": sum(d["exec"] for d in by.values()) / (conv_ms * 1e-3) / 1e12,
            "by_stage": {k: {"ms": v["ms"], "launches": v["n"], "alg_tflops": v["alg"] / (v["ms"] * 1e-3) / 1e12,
                             "frac_of_peak": v["alg"] / (v["ms"] * 1e-3) / 1e12 / pk["bf16_tflops"],
                             "frac_of_3pass_peak": 3 * v["alg"] / (v["ms"] * 1e-3) / 1e12 / pk["bf16_tflops"]} for k, v in by.items()},
            "note": "achieved = 405.5 GFLOP/face (SURVEY 8d: each output pixel once, conv_transpose at input resolution) x 16 faces / summed "
                    "CUDA-event durations of the convolution launches of one step; the 3-pass split issues 3 MMAs per product (so 1/3 of the "
                    "peak is the ceiling of this arithmetic: frac_of_3pass_peak) and the poly-phase up-convs execute 4x their algorithmic MACs "
                    "(executed_tflops counts those, not the split)"}


def generator_only(dev, steps):
    """BASELINE.json configs[1] (batch=16 synthesis from random regional style codes) for three mask families."""
    from e4s2024_b200 import engine as E
    from e4s2024_b200 import synth
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(SIZE, 512, 8, split_layer_idx=SPLIT, remaining_layer_idx=RL)
    synth.synth_module_weights(G, seed=2)
    G = G.to(dev).eval().requires_grad_(False)
    latent = synth.randn("bench.latent", (BATCH, K, 18, 512), 1).to(dev)
    out = {"workload": "configs[1]: batch=16 1024x1024 StyleGAN2 regional synthesis from random regional style codes, inputs resident",
           "alg_gflop_per_face": ALG_GFLOP["generator"], "masks": {}}
    for kind in ("blocky", "face", "noise"):
        mask = synth.onehot(synth.make_labels(kind, BATCH, K, 512, seed=1), K).to(dev)
        ms, _ = time_steps(lambda: G([latent], None, mask, input_is_latent=True, randomize_noise=False), steps, warmup=3)
        ctx = E.RegionCtx(mask, G._region_job_keys(), lazy=False, upz_keys=G._upz_keys())        # synchronous context: counts for the report
        jobs = {f"{h}x{w}{'_up' if up else ''}": round(rj.count / max(rj.tiles, 1), 3) for (h, w, up), rj in ctx.region_jobs.items()}
        rows = {f"{h}x{w}->{2 * h}x{2 * w}": round(uz.count / max(uz.cells_total, 1), 3) for (h, w), uz in ctx.upz.items()}
        out["masks"][kind] = {"value": BATCH / ms * 1e3, "unit": "faces/s", "ms_per_step": ms, "region_jobs_per_tile": jobs, "upconv_rows_per_cell": rows,
                              "alg_tflops": BATCH / ms * ALG_GFLOP["generator"]}
    out["mask_kinds"] = {"blocky": "random class per 32x32 cell of the 512^2 label map (tile-aligned at every resolution >= 64^2)",
                         "face": "procedural face (ellipses, curved boundaries; regions per 16x8 tile 1.4 / 1.8 / 2.7 / 4.9 at 256^2 / 128^2 / 64^2 / 32^2, "
                                 "the statistics of the reference's bundled CelebA-HQ masks, SURVEY B.6)",
                         "noise": "independent random class per pixel (adversarial: every tile sees all 12 regions)"}
    del G
    return out


def other_configs(hot, dev, steps=3):
    """configs[2] (batch=32 encoder -> regional styles -> synthesis round trip) and configs[3] (BiSeNet batch=64 at 512^2)."""
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import synth
    res = {}
    img = synth.smooth_image("cfg2.img", 32, SIZE, 17).to(dev)
    mask = synth.onehot(synth.make_labels("blocky", 32, K, 512, seed=17), K).to(dev)
    ms, _ = time_steps(lambda: hot.net(img, mask, randomize_noise=False), steps, warmup=2)
    res["net3_b32"] = {"workload": "configs[2]: batch=32 Net3.forward (encoder -> regional styles -> synthesis), blocky masks, inputs resident",
                       "value": 32 / ms * 1e3, "unit": "faces/s", "ms_per_step": ms,
                       "alg_tflops": 32 / ms * (ALG_GFLOP["encoder"] + ALG_GFLOP["mlps"] + ALG_GFLOP["generator"])}
    del img, mask
    torch.cuda.empty_cache()
    x = synth.randn("cfg3.x", (64, 512, 512, 8), 18)
    x[..., 3:] = 0
    x = x.to(dev)
    ms, _ = time_steps(lambda: hot.parser.seg.labels(x, (512, 512), None), steps, warmup=2)
    res["bisenet_b64"] = {"workload": "configs[3]: BiSeNet forward batch=64 512x512 -> 19-class label map (x8 upsample + argmax fused), inputs resident",
                          "value": 64 / ms * 1e3, "unit": "faces/s", "ms_per_step": ms, "alg_tflops": 64 / ms * ALG_GFLOP["parse"]}
    del x
    torch.cuda.empty_cache()
    # the reference pipelines call the path with ONE face at a time (face_swap_video_pipeline.py): latency of the whole hot path at batch 1
    u8 = synth.smooth_image_u8("lat.img", 1, SIZE, 19).to(dev)
    ms, _ = time_steps(lambda: hot.run_shard_u8(u8), max(steps, 5), warmup=3)
    res["latency_b1"] = {"workload": "the full swap hot path on one uint8 1024^2 face (batch 1, inputs resident): what a frame-by-frame pipeline sees",
                         "ms": ms, "value": 1e3 / ms, "unit": "faces/s"}
    try:                                          # the same call replayed as one CUDA graph (serving.GraphedSwapPath): no host launch cost
        from e4s2024_b200.serving import GraphedSwapPath
        gs = GraphedSwapPath(hot, batch=1, size=SIZE)
        ms_g, _ = time_steps(lambda: gs(u8), max(steps, 5), warmup=3)
        res["latency_b1"]["ms_cuda_graph"] = ms_g
        del gs
    except Exception as e:
        res["latency_b1"]["ms_cuda_graph"] = f"{type(e).__name__}: {e}"
    return res


def gpu_reference(sds, img_u8_d, batch=8, reps=2):
    """The 'existing Blackwell path' (BASELINE.md section 4.5): the reference ALGORITHM as plain torch ops on the same
    B200 -- K = 12 grouped cuDNN convolutions + mask multiply-adds per masked layer, materialised per-sample weights, separate
    blur / bias-act passes -- with TF32 off (fp32 parity mode) and on (cuDNN's default)."""
    dev = img_u8_d.device
    sd = {"net": {k: v.to(dev) for k, v in sds["net"].items()}, "seg": {k: v.to(dev) for k, v in sds["seg"].items()},
          "latent_avg": sds["latent_avg"].to(dev)}
    x = img_u8_d[:batch]
    out = {"workload": f"full path, batch={batch}, oracle/e4s_oracle.py ops on cuda (cuDNN {torch.backends.cudnn.version()}), inputs resident"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ms, _ = time_steps(lambda: oracle_chain(sd, x), reps, warmup=1)
            out[tag] = {"value": batch / ms * 1e3, "unit": "faces/s", "ms_per_step": ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del sd
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="e4s_b200", choices=["e4s_b200", "reference"])
    ap.add_argument("--engine", default=None, choices=[None, "tc", "f32"], help="conv engine override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs[1..3], gpu_reference and the stage / roofline pass")
    ap.add_argument("--dump-layers", default=None, help="write per-conv-launch timings (JSON lines) to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
    from e4s2024_b200 import synth
    from e4s2024_b200.serving import HostPipeline
    from e4s2024_b200.sharding import AsyncGather
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.engine:
        E.set_conv_engine(args.engine)
    args.warmup = max(args.warmup, 3)

    hot, sds = build_path(dev)
    img_h = synth.smooth_image_u8("swap.img", BATCH, SIZE, 13 + rank).pin_memory()        # what a pipeline holds: uint8 HWC
    img_d = img_h.to(dev)
    out_h = (torch.empty(BATCH, SIZE, SIZE, 3, dtype=torch.uint8).pin_memory(), torch.empty(BATCH, 512, 512, dtype=torch.uint8).pin_memory())
    # the path's only exchange: all-gather of the uint8 images + label maps, asynchronous and double-buffered
    gather = AsyncGather(world, [((BATCH, SIZE, SIZE, 3), torch.uint8), ((BATCH, 512, 512), torch.uint8)], dev) if world > 1 else None

    def step_resident():
        out_u8, labels = hot.run_shard_u8(img_d)
        if gather is not None:
            gather.submit([out_u8, labels])
        return out_u8, labels

    pipe = HostPipeline(hot.run_shard_u8, dev)

    def run_e2e(steps):
        for _ in range(steps):
            pipe.submit((img_h,), out_h)
        pipe.drain()

    def timed(fn, steps, whole=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
            if gather is not None:
                gather.wait()                            # every gather of the timed steps completes inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    step_resident()                                      # first pass also packs the weights
    torch.cuda.synchronize()
    n0 = L.launch_count()
    out_u8, labels = step_resident()
    launches_per_step = L.launch_count() - n0
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    run_e2e(2)
    ms_e2e = timed(run_e2e, args.steps, whole=True)
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (ms / 1e3)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)

    if os.environ.get("E4S_NCU"):          # one clean step for `ncu --profile-from-start off`
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        if gather is not None:
            gather.wait()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    line = {"metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3-pass hi/lo split on tcgen05 tensor cores: bf16 pairs in the encoder / generator, fp16 pairs in BiSeNet; fp32 accumulate)"
                     if E.conv_engine() == "tc" else "f32",
            "data": "synthetic", "config": bench_config(world, E.conv_engine()), "clocks": clocks,
            "gpu_launches": int(launches_per_step * args.steps),
            "e2e": {"value": e2e_value, "unit": "faces/s", "ms_per_step": ms_e2e / args.steps,
                    "note": "SwapHotPath.run_shard_u8 through serving.HostPipeline: pinned-host uint8 HWC images in (H2D every step), uint8 HWC images + "
                            "u8 label maps out (D2H every step), double-buffered on copy streams; timed from the first H2D to the last D2H complete"
                            + ("; every rank streams its own shard to / from host memory (no device all-gather on this leg)" if world > 1 else ""),
                    "h2d_bytes_per_step": int(img_h.numel()) * world, "d2h_bytes_per_step": int(out_h[0].numel() + out_h[1].numel()) * world},
            "alg_gflop_per_face": ALG_GFLOP_PER_FACE, "job_alg_tflops": value * ALG_GFLOP_PER_FACE / 1e3}
    if gather is not None:
        line["all_gather_bytes_per_rank_per_step"] = gather.bytes_per_rank()

    do_extras = rank == 0 and not args.no_extras
    if (rank == 0 and not args.no_extras) or (rank == 0 and args.dump_layers):
        stage_ms, prof = stage_breakdown(hot, img_d)
        tp = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")     # dram bytes of the same launches from one ncu capture of this command
        traffic = json.load(open(tp)).get("dram_bytes_per_step") if os.path.exists(tp) else None
        line["roofline"] = roofline_from(prof, BATCH, ms / args.steps, traffic)
        if line["roofline"] is not None:
            line["roofline"]["traffic_source"] = "profiles/r2_ncu_traffic.json (ncu capture of this command, sum over the step's convolution launches)" if traffic else None
        line["stage_ms"] = stage_ms
        if args.dump_layers and rank == 0:
            with open(args.dump_layers, "w") as f:
                for r in prof:
                    ms_l = r["ev"][0].elapsed_time(r["ev"][1])
                    f.write(json.dumps({"stage": r["stage"], "engine": r["engine"], "fmt": r["fmt"], "m": r["m"], "k": r["k"], "n": r["n"], "cin": r["cin"],
                                        "hout": r["hout"], "kh": r["kh"], "stride": r["stride"], "up2": r["up2"], "ms": round(ms_l, 4),
                                        "alg_tflops": round(r["alg_flops"] / ms_l / 1e9, 2), "exec_tflops": round(r["exec_flops"] / ms_l / 1e9, 2),
                                        "io_gbs": round(r["bytes"] / ms_l / 1e6, 1)}) + "\n")
    if do_extras and world == 1:
        del pipe
        torch.cuda.empty_cache()
        for key, fn in (("configs", lambda: other_configs(hot, dev)), ("gpu_reference", lambda: gpu_reference(sds, img_d))):
            try:
                r = fn()
                line.update(r) if key == "configs" else line.__setitem__(key, r)
            except Exception as e:                      # a secondary line must never take the headline down with it
                line[key] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    if rank == 0:
        gpu_lab1 = labels[:1].cpu().numpy()
        x01_1, x_1 = L.im2tensor(img_d[:1])
        fimg = hot.run_shard(x_1, img01=x01_1)[0].cpu()          # fp32 image of face 0 (bit-identical to its slot in the batch)
        if world == 1 and not args.no_extras:
            try:
                line["parity"] = label_parity(sds, img_h[:1], gpu_lab1)
            except Exception as e:
                line["parity"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline and world == 1:
            cpu, ref_img, ref_lab = cpu_baseline(sds, img_h[:1], gpu_lab1)
            line["cpu_baseline"] = cpu
            line.setdefault("parity", {})
            line["parity"].update({"image_max_abs_diff_vs_cpu_oracle": float((fimg - ref_img).abs().max()), "image_abs_range": float(ref_img.abs().max()),
                                   "image_tolerance": 1e-3, "faces_compared": 1})
        else:
            line["cpu_baseline"] = None
    if do_extras and world == 1:
        del hot
        torch.cuda.empty_cache()
        try:
            line["generator_only"] = generator_only(dev, min(args.steps, 10))
        except Exception as e:
            line["generator_only"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

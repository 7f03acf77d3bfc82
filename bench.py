#!/usr/bin/env python
"""bench.py -- 1024^2 faces/sec of the E4S hot path on N B200s (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- batch=16 per GPU, 1024x1024 StyleGAN2 regional
synthesis (Generator(1024, rl=13, split=5), K=12 regions) from random regional style codes, blocky one-hot
masks, fixed noise buffers, synthetic weights.  A "step" = one Generator.forward over the batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this framework (CUDA, libe4s_b200.so)
  python bench.py --impl reference ...                         # the reference algorithm on the host CPU cores

value  : whole-job faces/s with inputs resident in HBM (CUDA events, max over ranks).
e2e    : same metric through the public nn.Module API with HOST (pinned) latent+mask copied in and the
         images copied back to pinned host memory inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "1024^2 faces/sec (StyleGAN2 regional synthesis, Generator 1024 K=12 rl=13)"
BATCH, SIZE, K, RL, SPLIT = 16, 1024, 12, 13, 5
ALG_GFLOP_PER_FACE = 148.52          # SURVEY.md section 8(d) per-layer table (each output pixel once)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "hbm_gbs": d["hbm_gbs"],
                "source": "MEASURED_PEAKS.json (sustained bf16, copy bandwidth)"}
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


def bench_config(world, engine="tc"):
    """`config` of the JSON line: the same dict for both arms (the reference arm adds what its bounded sample was)."""
    return {"workload": "configs[1]: batch=16/GPU 1024x1024 StyleGAN2 regional synthesis from random regional style codes",
            "batch_per_gpu": BATCH, "global_batch": BATCH * world, "size": SIZE, "regions": K, "remaining_layer_idx": RL,
            "masks": "blocky one-hot 32x32 cells @512^2", "noise": "registered buffers (randomize_noise=False)",
            "parallelism": f"batch-sharded x{world}" + (" + NCCL all_gather of images" if world > 1 else ""),
            "conv_engine": engine, "l2": "working set (>2 GB activations per step) exceeds the 126 MB L2"}


def swap_path_line(dev, steps=3):
    """BASELINE.json's metric names the whole swap hot path (parsing + encoder + synthesis, configs[4] per GPU shard):
    FaceParser.parse_batch -> one-hot -> Net3.forward at 16 faces, inputs resident, CUDA events, stage by stage."""
    from e4s2024_b200 import _lib as L, synth
    from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
    from e4s2024_b200.networks import Net3
    import types
    # the fields Net3 reads from the reference's option object (options/our_swap_face_pipeline_options.py:12-18,50)
    opts = types.SimpleNamespace(fsencoder_type="psp", remaining_layer_idx=RL, num_seg_cls=K, out_size=SIZE, train_G=False,
                                 start_from_latent_avg=True, learn_in_w=False)
    net = Net3(opts)
    synth.synth_module_weights(net, seed=9)
    net = net.to(dev)
    net.latent_avg = synth.randn("net3.latent_avg", (18, 512), 9, 0.1).to(dev)
    parser = FaceParser(seg_ckpt=None, size=SIZE, device=str(dev))
    synth.synth_module_weights(parser.seg, seed=10)
    parser.seg.to(dev)
    img = synth.smooth_image("swap.img", BATCH, SIZE, 13).to(dev)
    img01 = (img + 1) / 2

    def t(fn):
        r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, r

    ms_parse, lab = t(lambda: parser.parse_batch(img01))
    ms_onehot, mask = t(lambda: L.labels_to_onehot(lab, K))
    ms_enc, (vec, _) = t(lambda: net.get_style_vectors(img, mask))
    ms_codes, codes = t(lambda: net.cal_style_codes(vec))
    ms_gen, _ = t(lambda: net.gen_img(None, codes, mask, randomize_noise=False))
    ms_all, _ = t(lambda: net(img, L.labels_to_onehot(parser.parse_batch(img01), K), randomize_noise=False))
    # opt-in fast parse (bf16x3 tensor-core BiSeNet): reported beside the exact mode, never instead of it
    from e4s2024_b200.face_parsing import resnet as _rn
    _rn.set_bisenet_engine("tc")
    ms_parse_tc, lab_tc = t(lambda: parser.parse_batch(img01))
    ms_all_tc, _ = t(lambda: net(img, L.labels_to_onehot(parser.parse_batch(img01), K), randomize_noise=False))
    _rn.set_bisenet_engine("f32")
    differing = int((lab_tc != lab).sum())
    del net, parser
    return {"workload": "configs[4] per-GPU shard: 16 faces, bicubic 1024->512 + BiSeNet + argmax/LUT -> one-hot -> Net3 (encoder, 12 MLPs, generator)",
            "value": BATCH / ms_all * 1e3, "unit": "faces/s", "ms_per_step": ms_all,
            "stage_ms": {"parse": ms_parse, "onehot": ms_onehot, "encoder": ms_enc, "mlps": ms_codes, "generator": ms_gen},
            "alg_gflop_per_face": 405.5, "note": "BiSeNet runs on the exact-fp32 CUDA-core engine (bit-exact label maps), the rest on tcgen05",
            "fast_parse_opt_in": {"value": BATCH / ms_all_tc * 1e3, "unit": "faces/s", "ms_per_step": ms_all_tc, "parse_ms": ms_parse_tc,
                                  "labels_differing_from_exact_mode": differing, "labels_total": int(lab.numel()),
                                  "note": "E4S_BISENET_ENGINE=tc: BiSeNet on the bf16x3 tensor-core engine; does not meet the bit-exact label bar, not the default"}}


def make_generator_inputs(batch, seed=1):
    from e4s2024_b200 import synth
    latent = synth.randn("bench.latent", (batch, K, 18, 512), seed)
    labels = synth.blocky_labels(batch, K, 512, cells=32, seed=seed)
    return latent, synth.onehot(labels, K)


def build_generator(device):
    from e4s2024_b200 import synth
    from e4s2024_b200.stylegan2.model import Generator
    G = Generator(SIZE, 512, 8, split_layer_idx=SPLIT, remaining_layer_idx=RL)
    synth.synth_module_weights(G, seed=2)
    return G.to(device).eval()


def thread_candidates():
    """torch/oneDNN does not scale to every hardware thread on big hosts (128 threads were 6x slower than 8 on the
    first box), so the CPU arm tries a few thread counts and keeps the fastest; `cores` reports the one used."""
    n = os.cpu_count() or 1
    return sorted({n, min(n, 64), min(n, 32), min(n, 16)}, reverse=True)


def cpu_baseline(faces=1):
    """The oracle (CPU restatement of the reference algorithm: K=12 grouped convs per masked layer) on the host cores."""
    from e4s2024_b200 import synth
    from e4s2024_b200.stylegan2.model import generator_state_shapes
    from oracle import e4s_oracle as orc
    sd = synth.fill_state_dict(generator_state_shapes(SIZE, split_layer_idx=SPLIT, remaining_layer_idx=RL), seed=2)
    latent, mask = make_generator_inputs(faces)
    best, best_t, tried = None, None, {}
    with torch.no_grad():
        for t in thread_candidates():
            torch.set_num_threads(t)
            t0 = time.perf_counter()
            orc.generator_forward(sd, SIZE, latent, mask, split_layer_idx=SPLIT, remaining_layer_idx=RL)
            dt = time.perf_counter() - t0
            tried[t] = round(dt, 2)
            if best is None or dt < best:
                best, best_t = dt, t
            if dt > 45:                       # keep the whole bench within minutes
                continue
    return {"value": faces / best, "unit": "faces/s", "cores": best_t, "kind": "port",
            "sample": f"{faces} face, Generator 1024^2 K=12 rl=13, oracle/e4s_oracle.py fp32; seconds per thread count {tried}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = bench_config(max(int(os.environ.get("WORLD_SIZE", "1")), 1), engine="cpu")
    cfg["sample"] = "1 face per step (bounded sample of the 16-face batch), CPU reference algorithm, all host threads that help"
    from e4s2024_b200 import synth
    from e4s2024_b200.stylegan2.model import generator_state_shapes
    from oracle import e4s_oracle as orc
    sd = synth.fill_state_dict(generator_state_shapes(SIZE, split_layer_idx=SPLIT, remaining_layer_idx=RL), seed=2)
    latent, mask = make_generator_inputs(1)
    step = lambda: orc.generator_forward(sd, SIZE, latent, mask, split_layer_idx=SPLIT, remaining_layer_idx=RL)
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
                                       "sample": "1 face per step, oracle/e4s_oracle.py generator_forward"},
                      "e2e": {"value": v, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="e4s_b200", choices=["e4s_b200", "reference"])
    ap.add_argument("--engine", default=None, choices=[None, "tc", "f32"], help="conv engine override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-swap-path", action="store_true", help="skip the full swap-path (parser + encoder + generator) timing")
    ap.add_argument("--graph", action="store_true", help="replay one CUDA graph per forward (serving.GraphedGenerator) instead of launching "
                    "every kernel from the host; measured equal on one B200 (the step is GPU-bound), so the eager path stays the default")
    ap.add_argument("--dump-layers", default=None, help="write per-conv-launch timings (JSON lines) to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from e4s2024_b200 import _lib as L
    from e4s2024_b200 import engine as E
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

    G = build_generator(dev)
    latent_h, mask_h = make_generator_inputs(BATCH, seed=1 + rank)
    latent_h, mask_h = latent_h.pin_memory(), mask_h.pin_memory()
    latent_d, mask_d = latent_h.to(dev), mask_h.to(dev)
    # the path's only exchange: all-gather of the output images, double-buffered and asynchronous so that the gather of
    # step i rides under the kernels of step i+1 (NVLink/NVSwitch traffic, no dependence on the next step's inputs)
    gathered = [torch.empty(world * BATCH, 3, SIZE, SIZE, device=dev) for _ in range(2)] if world > 1 else None
    inflight = []
    out_h = torch.empty(BATCH, SIZE, SIZE, 3, dtype=torch.uint8).pin_memory()     # what tensor2im hands the host: uint8 HWC images

    # One CUDA graph per forward (serving.GraphedGenerator): the ~40 launches of a step replay as one; the eager path stays for
    # the per-launch timing pass and --no-graph.
    from e4s2024_b200.serving import GraphedGenerator
    G([latent_d], None, mask_d, input_is_latent=True, randomize_noise=False)          # first forward also packs the weights
    n0 = L.launch_count()
    G([latent_d], None, mask_d, input_is_latent=True, randomize_noise=False)
    launches_per_forward = L.launch_count() - n0
    gg = None if not args.graph else GraphedGenerator(G, BATCH, K, (mask_d.shape[2], mask_d.shape[3]), device=dev)

    def gen(lat, msk):
        if gg is None:
            return G([lat], None, msk, input_is_latent=True, randomize_noise=False)[0]
        return gg(lat, msk)

    def step_resident():
        img = gen(latent_d, mask_d)
        if world > 1 and gg is not None:
            img = img.clone()                            # the graph's static output is overwritten by the next replay
        if world > 1:
            if len(inflight) == 2:                       # the buffer about to be reused: its gather must have completed
                inflight.pop(0).wait()
            buf = gathered[step_resident.n % 2]
            step_resident.n += 1
            inflight.append(dist.all_gather_into_tensor(buf, img, async_op=True))
        return img

    step_resident.n = 0

    def finish_gathers():
        while inflight:
            inflight.pop(0).wait()

    # end to end: every step copies its inputs (latent + one-hot mask) from pinned host memory and its images back;
    # the copies of step i+1 / i-1 overlap the kernels of step i (e4s2024_b200/serving.py), as a serving loop would
    # The mask crosses PCIe as the u8 label map the pipelines hold on the host and becomes one-hot on the device
    # (utils.torch_utils.labelMap2OneHot, as in the reference's own flow: label map -> one-hot -> Generator).
    from e4s2024_b200.serving import HostPipeline
    # The images leave as the uint8 HWC arrays the pipelines build right after the generator (utils.torch_utils.tensor2im,
    # face_swap_video_pipeline.py: `tensor2im(swapped_face_image[0])`), converted on the device with identical arithmetic.
    from e4s2024_b200.utils.torch_utils import labelMap2OneHot, tensor2im_batch
    labels_h = mask_h.argmax(1, keepdim=True).to(torch.uint8).pin_memory()          # [B,1,512,512]

    def e2e_fn(lat, lab):
        return tensor2im_batch(gen(lat, labelMap2OneHot(lab, K)))

    pipe = HostPipeline(e2e_fn, dev)

    def run_e2e(steps):
        for _ in range(steps):
            pipe.submit((latent_h, labels_h), out_h)
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
            if world > 1:
                finish_gathers()                         # every gather of the timed steps completes inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    launches = args.steps * launches_per_forward         # kernels executed per step (a graph replay runs the same nodes)
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (ms / 1e3)

    run_e2e(2)
    ms_e2e = timed(run_e2e, args.steps, whole=True)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)

    if os.environ.get("E4S_NCU"):          # one clean step for `ncu --profile-from-start off`
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---- per-launch timing pass for the roofline of the dominant (convolution) kernel -----------------
    E.PROFILE = []
    step_resident()
    torch.cuda.synchronize()
    prof, E.PROFILE = E.PROFILE, None
    if args.dump_layers and rank == 0:
        with open(args.dump_layers, "w") as f:
            for r in prof:
                ms_l = r["ev"][0].elapsed_time(r["ev"][1])
                f.write(json.dumps({"engine": r["engine"], "m": r["m"], "k": r["k"], "n": r["n"], "up2": r["up2"], "ms": round(ms_l, 4),
                                    "exec_tflops": round(r["exec_flops"] / ms_l / 1e9, 2),
                                    "io_gbs": round(r["bytes"] / ms_l / 1e6, 1)}) + "\n")
    by = {}
    for r in prof:
        d = by.setdefault(r["engine"], {"ms": 0.0, "alg": 0.0, "exec": 0.0, "n": 0})
        d["ms"] += r["ev"][0].elapsed_time(r["ev"][1])
        d["alg"] += r["alg_flops"]
        d["exec"] += r["exec_flops"]
        d["n"] += 1
    pk = peaks()
    dom = max(by, key=lambda k: by[k]["ms"]) if by else None
    roof = None
    if dom:
        d = by[dom]
        ach = d["alg"] / (d["ms"] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")       # dram bytes of the same 17 launches, one ncu --set full capture
        if dom == "tc" and os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_step")
        roof = {"bound": "tensor", "kernel": "conv_tc_{halo,wide,gather} kernels (tcgen05 bf16x3), the 17 convolution launches of one step" if dom == "tc" else "conv_igemm_f32_kernel (CUDA-core fp32)",
                "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": traffic,
                "peak_source": pk["source"], "launches": d["n"], "kernel_ms_per_step": d["ms"],
                "executed_tflops": d["exec"] / (d["ms"] * 1e-3) / 1e12,
                "note": "achieved = algorithmic conv FLOPs (each output pixel once, conv_transpose at input resolution) / summed "
                        "CUDA-event durations of the conv launches of one step; the bf16x3 split issues 3 MMAs per product and "
                        "the poly-phase up-convs execute 4x the algorithmic MACs (executed_tflops counts the latter, not the split)",
                "by_engine": {k: {"ms": v["ms"], "launches": v["n"], "alg_tflops": v["alg"] / (v["ms"] * 1e-3) / 1e12} for k, v in by.items()}}

    used_graph = gg is not None
    swap = None
    if rank == 0 and world == 1 and not args.no_swap_path:
        del G, pipe, gg
        torch.cuda.empty_cache()
        try:
            swap = swap_path_line(dev)
        except Exception as e:                          # the secondary line must never take the headline down with it
            swap = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline()
        line = {"metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (bf16x3 split on tensor cores, fp32 accumulate)" if E.conv_engine() == "tc" else "f32",
                "data": "synthetic",
                "config": bench_config(world, E.conv_engine()),
                "clocks": clocks, "gpu_launches": int(launches), "cuda_graph": used_graph,
                "e2e": {"value": e2e_value, "unit": "faces/s", "ms_per_step": ms_e2e / args.steps,
                        "note": "labelMap2OneHot + Generator.forward + tensor2im_batch through HostPipeline: pinned-host H2D of every step's latent + u8 "
                                "label map, D2H of its uint8 HWC images (the reference pipelines' tensor2im output), double-buffered on copy "
                                "streams (timed region = first H2D to last D2H complete)",
                        "h2d_bytes_per_step": int(latent_h.numel() * 4 + labels_h.numel()), "d2h_bytes_per_step": int(out_h.numel())},
                "roofline": roof, "cpu_baseline": cpu, "swap_path": swap,
                "alg_gflop_per_face": ALG_GFLOP_PER_FACE,
                "job_alg_tflops": value * ALG_GFLOP_PER_FACE / 1e3}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

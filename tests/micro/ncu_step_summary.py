"""Summarise `ncu --page raw --csv` of one bench step (per launch rows) into (a) a per-kernel table (launches, total ms, share of the
step, DRAM bytes, time-weighted tensor-pipe active %) and (b) the DRAM traffic of the convolution launches (roofline.traffic).
    python tests/micro/ncu_step_summary.py gpurun_out/ncu_step_r2.csv profiles/r2_ncu_step_summary.tsv profiles/r2_ncu_traffic.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
ix = {n: h.index(n) for n in h}
def f(r, n, default=0.0):
    try:
        return float(r[ix[n]].replace(",", ""))
    except Exception:
        return default
def unit_scale(n, want):
    u = units[ix[n]].lower() if n in ix else ""
    tbl = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3} if want == "ms" else {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    return tbl.get(u, 1.0)
ts, bs_r, bs_w = unit_scale("gpu__time_duration.sum", "ms"), unit_scale("dram__bytes_read.sum", "b"), unit_scale("dram__bytes_write.sum", "b")
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
agg = {}
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("e4s::", "")
    d = agg.setdefault(name, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "tp_ms": 0.0, "regs": r[ix["launch__registers_per_thread"]] if "launch__registers_per_thread" in ix else "-"})
    ms = f(r, "gpu__time_duration.sum") * ts
    d["n"] += 1; d["ms"] += ms; d["rd"] += f(r, "dram__bytes_read.sum") * bs_r; d["wr"] += f(r, "dram__bytes_write.sum") * bs_w
    d["tp_ms"] += f(r, TP) * ms
tot = sum(d["ms"] for d in agg.values())
with open(sys.argv[2], "w") as o:
    o.write("kernel\tlaunches\ttotal_ms\tshare_of_step\tdram_read_GB\tdram_write_GB\tdram_GBs\ttensor_pipe_active_pct_time_weighted\tregs\n")
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        o.write(f"{name}\t{d['n']}\t{d['ms']:.4f}\t{d['ms'] / tot:.4f}\t{d['rd'] / 1e9:.4f}\t{d['wr'] / 1e9:.4f}\t{(d['rd'] + d['wr']) / max(d['ms'], 1e-9) / 1e6:.1f}\t{d['tp_ms'] / max(d['ms'], 1e-9):.2f}\t{d['regs']}\n")
    o.write(f"TOTAL\t{sum(d['n'] for d in agg.values())}\t{tot:.4f}\t1.0\t{sum(d['rd'] for d in agg.values()) / 1e9:.4f}\t{sum(d['wr'] for d in agg.values()) / 1e9:.4f}\t-\t-\t-\n")
conv = {k: v for k, v in agg.items() if k.startswith("conv_tc") or k.startswith("conv_igemm")}
json.dump({"dram_bytes_per_step": sum(v["rd"] + v["wr"] for v in conv.values()), "dram_read_bytes": sum(v["rd"] for v in conv.values()),
           "dram_write_bytes": sum(v["wr"] for v in conv.values()), "launches": sum(v["n"] for v in conv.values()),
           "conv_ms_under_ncu": sum(v["ms"] for v in conv.values()), "step_ms_under_ncu": tot,
           "conv_share_of_step_under_ncu": sum(v["ms"] for v in conv.values()) / tot,
           "tensor_pipe_active_pct_time_weighted": sum(v["tp_ms"] for v in conv.values()) / max(sum(v["ms"] for v in conv.values()), 1e-9),
           "source": "ncu --clock-control none --metrics dram__bytes_{read,write}.sum,... of one full-path bench step (python bench.py, E4S_NCU=1), every conv_tc_* / conv_igemm launch, B=16"},
          open(sys.argv[3], "w"), indent=1)
print(open(sys.argv[2]).read())

"""Summarise an `ncu --page raw --csv` dump (one row per profiled launch) into the columns the roofline needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return h.index(name) if name in h else None
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "dur"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct_elapsed"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_rt_pct"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("launch__registers_per_thread", "regs")]
print("\t".join(f"{n}[{units[col(c)]}]" if col(c) is not None else n for c, n in cols))
for r in data:
    out = []
    for c, n in cols:
        i = col(c)
        v = r[i] if i is not None else "-"
        if n == "kernel":
            v = v.split("(")[0][-40:]
        out.append(v)
    print("\t".join(out))

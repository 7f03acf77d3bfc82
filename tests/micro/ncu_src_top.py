"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples with their dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# the file may hold several kernels: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h = b["rows"][0]
    si = h.index("# Samples"); ai = h.index("Address"); so = h.index("Source"); ie = h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = [r for r in b["rows"][1:] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data)
    print(f"== {b['name'][:90]}  total samples {tot}, {len(data)} instrs")
    agg = {}
    for r in data:
        for i, c in stall_cols:
            v = int(r[i] or 0)
            agg[c] = agg.get(c, 0) + v
    print("  by reason:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for n, r in enumerate(sorted(data, key=lambda r: -int(r[si]))[:top]):
        reasons = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:2]
        idx = data.index(r)
        print(f"  {100*int(r[si])/max(tot,1):5.1f}%  #{idx:5d} exec {r[ie]:>9s}  {r[so].strip()[:70]:70s} {reasons}")

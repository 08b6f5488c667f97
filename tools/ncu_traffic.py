"""dev tool: per-launch DRAM traffic of one kernel from an ncu launch list (CSV of
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`), and the per-kernel shares.
usage: python tools/ncu_traffic.py launches.csv [kernel-substring]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ksub = sys.argv[2] if len(sys.argv) > 2 else "k_wf_trace"
by = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = by.setdefault(d["ID"], {"name": d["Kernel Name"]})
    unit, val = d["Metric Unit"], float(d["Metric Value"].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    e[d["Metric Name"]] = val * scale
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in by.values():
    short = e["name"].split("(")[0].replace("void ", "")
    t = tot[short]
    t[0] += 1
    t[1] += e.get("gpu__time_duration.sum", 0.0)
    t[2] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
all_ns = sum(t[1] for t in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'DRAM GB':>9s} {'MB/launch':>10s}")
for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {t[0]:8d} {t[1] / 1e6:10.2f} {100 * t[1] / all_ns:6.1f}% {t[2] / 1e9:9.2f} {t[2] / 1e6 / t[0]:10.1f}")
sel = [t for k, t in tot.items() if ksub in k]
n = sum(t[0] for t in sel); b = sum(t[2] for t in sel)
print(f"\n{ksub}: {n} launches, {b / 1e9:.2f} GB DRAM, {b / max(1, n):.4e} bytes per launch")

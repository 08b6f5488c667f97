"""dev tool: markdown summaries of ncu outputs for profiles/.

  python tools/ncu_summary.py launches  <launches.csv>             per-kernel totals, share of the run, DRAM bytes
  python tools/ncu_summary.py metrics   <report.ncu-rep> [index]   key metrics of one captured launch
"""
import collections
import csv
import io
import subprocess
import sys

KEY = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per issued instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (conversion) pipe busy"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe (LSU wavefronts) busy"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (per issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("sass__inst_executed_local_loads", "local loads (stack + spills)"),
    ("sass__inst_executed_local_stores", "local stores"),
]


def to_us(value, unit):
    v = float(value.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        if "cub::" in name:
            name = "cub::" + name.split("cub::")[1].split("<")[0]
        d = per.setdefault((row["ID"], name), {})
        d[row["Metric Name"]] = (row["Metric Value"], row["Metric Unit"])
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, name), m in per.items():
        t = tot[name]
        t[0] += 1
        if "gpu__time_duration.sum" in m:
            t[1] += to_us(*m["gpu__time_duration.sum"])
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in m:
                t[2] += to_bytes(*m[k])
    total = sum(t[1] for t in tot.values())
    print("| kernel | launches | total ms | share | DRAM GB (read+write) |")
    print("|---|---|---|---|---|")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {t[0]} | {t[1] / 1e3:.3f} | {100 * t[1] / total:.1f} % | {t[2] / 1e9:.2f} |")
    print(f"| all | {sum(t[0] for t in tot.values())} | {total / 1e3:.3f} | 100 % | {sum(t[2] for t in tot.values()) / 1e9:.2f} |")


def metrics(path, index=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2 + index]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"kernel: `{vals[col['Kernel Name']]}`\n")
    print("| metric | value | unit | ncu name |")
    print("|---|---|---|---|")
    for key, label in KEY:
        if key in col:
            v = vals[col[key]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            print(f"| {label} | {v} | {units[col[key]]} | `{key}` |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        metrics(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)

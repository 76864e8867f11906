#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small tracked summaries under profiles/.

  launches CSV (`ncu --metrics gpu__time_duration.sum ...`)  -> profiles/<tag>_launches_summary.csv  (per-kernel share)
  *.ncu-rep (`ncu --set full ...`)                            -> profiles/<tag>_ncu_<name>.md       (key metrics + hot SASS)
Usage: tools/summarise_profiles.py <tag> [--launches gpurun_out/launches.csv] [--rep name=path ...]
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("dicow::<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def launches(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1000.0)
            rows.append((short(r["Kernel Name"]), r["Grid Size"], us))
    agg = OrderedDict()
    for k, g, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary: {len(rows)} launches, total {total / 1000.0:.2f} ms "
                "(per-launch times under ncu are cold-cache / serialised: compare SHARES, not absolutes)\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{us:.1f},{us / n:.1f},{us / total:.3f}\n")
    print(open(out).read())


def rep(name, path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary: {name} ({path.split('/')[-1]})\n\n")
        for r in rows[2:]:
            kn = short(r[hdr.index("Kernel Name")])
            f.write(f"## {kn}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in KEY_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            f.write("\n")
        hot = subprocess.run([sys.executable, "tools/ncu_hot.py", path, "12"], capture_output=True, text=True).stdout
        f.write("## stall / instruction mix and hottest SASS lines (source page)\n\n```\n" + hot + "```\n")
    print("wrote", out)


if __name__ == "__main__":
    tag = sys.argv[1]
    args = sys.argv[2:]
    i = 0
    while i < len(args):
        if args[i] == "--launches":
            launches(args[i + 1], f"profiles/{tag}_launches_summary.csv")
            i += 2
        elif args[i] == "--rep":
            n, p = args[i + 1].split("=", 1)
            rep(n, p, f"profiles/{tag}_ncu_{n}.md")
            i += 2
        else:
            raise SystemExit(f"unknown arg {args[i]}")

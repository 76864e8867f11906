#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of a report: per kernel, the top-N SASS lines by stall samples and the
executed-instruction mix.  Usage: tools/ncu_hot.py report.ncu-rep [N] [kernel-regex]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
seen = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        i = j
        short = re.sub(r"\(CUtensorMap.*", "", name)
        if (pat and not pat.search(name)) or short in seen:
            continue
        seen.add(short)
        c = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[c["# Samples"]]) for r in body)
        inst = sum(int(r[c["Instructions Executed"]]) for r in body)
        print(f"== {short}: {tot} samples, {inst} warp-instructions executed")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[c[h]]) for r in body) for h in stall_cols}
        print("   stall mix: " + ", ".join(f"{h[6:]} {100*v/max(1,tot):.0f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        mix = {}
        for r in body:
            op = r[c["Source"]].split()[0] if not r[c["Source"]].strip().startswith("@") else r[c["Source"]].split()[1]
            op = op.split(".")[0]
            mix[op] = mix.get(op, 0) + int(r[c["Instructions Executed"]])
        print("   inst mix: " + ", ".join(f"{k} {100*v/max(1,inst):.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
        for r in sorted(body, key=lambda r: -int(r[c["# Samples"]]))[:N]:
            s = int(r[c["# Samples"]])
            top = max(stall_cols, key=lambda h: int(r[c[h]]))
            print(f"   {100*s/max(1,tot):5.1f}%  {r[c['Source']].strip()[:70]:70s} {top[6:]}")
    else:
        i += 1

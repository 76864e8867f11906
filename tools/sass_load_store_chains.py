#!/usr/bin/env python
"""Static check for the pattern that cost the decode step 20 % until round 2's third session: a copy / staging loop whose
global loads are each followed by a dependent store, so that the loop runs one memory round trip per iteration instead of
keeping its loads in flight (`ld.global` then `st.shared` of a staged slab compiled to
LDG, STS, LDG, STS, ...).  Needs no GPU: reads `cuobjdump -sass` of the object files under ts-asr-whisper_b200/build/.

For every loop (a backward branch) of at most 400 instructions it prints the order of loads (L) and stores (S) when loads and
stores alternate at least three times.  A hit is a CANDIDATE: a loop that is bandwidth-bound with many warps in flight (the
AdamW kernel) is fine; a loop executed by a handful of latency-bound CTAs (beam_select_kernel: one CTA per utterance) is not.

Usage: python tools/sass_load_store_chains.py [object files ...]"""
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
objs = sys.argv[1:] or sorted(glob.glob(os.path.join(HERE, "..", "ts-asr-whisper_b200", "build", "*.o")))
for obj in objs:
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0]
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        index = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.U)?\s+(?:[!U]*P\d+,\s*)?0x([0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= a or tgt not in index or i - index[tgt] > 400:
                continue
            seq = ""
            for _, bt in ins[index[tgt]:i + 1]:
                op = bt.split()[1] if bt.startswith("@") else bt.split()[0]
                if op.startswith(("LDG", "LD.")):
                    seq += "L"
                elif op.startswith(("STS", "STG", "ST.")):
                    seq += "S"
            runs = re.sub(r"L+", "L", re.sub(r"S+", "S", seq))
            if seq.count("L") >= 3 and runs.count("LS") >= 3:
                mm = re.search(r"_cu_[0-9a-f]{8}(\d+)", name)  # Itanium mangling: <length><identifier>
                short = name[mm.end():mm.end() + int(mm.group(1))] if mm else name[:40]
                print(f"{os.path.basename(obj):16s} {short:28s} loop@{tgt:#x} ({i - index[tgt] + 1} instr)  {seq[:48]}")

#!/usr/bin/env python3
"""Attribute an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) to CUDA source lines
using nvdisasm line info of the cubin.  Usage: ncu_lines.py src.csv kernel.cubin kernel_name [topN]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the function's text: remember the current line marker for each instruction offset
line_of = {}
cur = None
infn = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = kname in ln
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = None
agg = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        base = None
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    inst = int(r[hdr.index("Instructions Executed")] or 0)
    samp = int(r[hdr.index("# Samples")] or 0)
    key = line_of.get(off, ((None, 0), ""))[0]
    agg[key][0] += inst
    agg[key][1] += samp
    agg[key][2] += 1
    tot += inst
    break_after_first_kernel = False
print(f"total warp instructions: {tot}")
src_cache = {}
def src_line(key):
    if not key or not key[0]:
        return ""
    import glob
    for p in glob.glob("/root/repo/dsp_audioreclabs_b200/csrc/" + key[0]):
        if p not in src_cache:
            src_cache[p] = open(p).read().splitlines()
        L = src_cache[p]
        return L[key[1] - 1].strip()[:90] if key[1] - 1 < len(L) else ""
    return ""
tots = sum(v[1] for v in agg.values())
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100.0 * v[0] / tot:6.2f}% inst {100.0 * v[1] / max(tots,1):6.2f}% samp  n_sass={v[2]:4d}  {key}  {src_line(key)}")

# optional phase table: PHASES="name:file:lo-hi,..." in the environment
import os
ph = os.environ.get("PHASES")
if ph:
    print("\nphase aggregates:")
    for spec in ph.split(","):
        name, f, rng = spec.split(":")
        lo, hi = (int(v) for v in rng.split("-"))
        i = sum(v[0] for k, v in agg.items() if k and k[0] == f and lo <= k[1] <= hi)
        s = sum(v[1] for k, v in agg.items() if k and k[0] == f and lo <= k[1] <= hi)
        print(f"  {name:28s} {100.0 * i / tot:6.2f}% inst  {100.0 * s / max(tots, 1):6.2f}% samples   {i / 8.8208e8:6.3f} warp-inst/sample")
    other_i = sum(v[0] for k, v in agg.items() if not k or k[0] not in ("frontend_pcm.cu", "common.cuh"))
    other_s = sum(v[1] for k, v in agg.items() if not k or k[0] not in ("frontend_pcm.cu", "common.cuh"))
    print(f"  {'(headers/intrinsics/none)':28s} {100.0 * other_i / tot:6.2f}% inst  {100.0 * other_s / max(tots, 1):6.2f}% samples")

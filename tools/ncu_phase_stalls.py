#!/usr/bin/env python3
"""Where the warps of a kernel spend their time, by phase: sums the sampled warp states of an ncu source page
(`ncu -i X.ncu-rep --page source --csv`) over CUDA source line ranges.  A sample is one warp observed at one
instruction; `samples / total` is the share of warp-time, the stall columns say what the warp was waiting for.
PHASES="name:lo-hi,..." are lines of the kernel's main source file; HELPER_BELOW as in ncu_cost.py.
Usage: ncu_phase_stalls.py src.csv kernel.cubin kernel_name main_source.cu"""
import csv, os, re, subprocess, sys
from collections import defaultdict
src_csv, cubin, kname, mainsrc = sys.argv[1:5]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}; stack = None; infn = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln): infn = kname in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = m.group(1).split("/")[-1], int(m.group(2)), m.group(3)
        stack = [(f, l)] + [(a.split("/")[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: line_of[int(m.group(1), 16)] = stack
rows = list(csv.reader(open(src_csv)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
phases = []
for spec in os.environ.get("PHASES", "").split(","):
    if spec:
        name, rng = spec.split(":"); lo, hi = (int(v) for v in rng.split("-")); phases.append((name, lo, hi))
helper_below = int(os.environ.get("HELPER_BELOW", "0"))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(float))
cur = "other"
for r in sorted(data, key=lambda r: int(r[0], 16)):
    off = int(r[0], 16) - base
    frames = line_of.get(off) or []
    ml = next((l for f, l in frames if f == mainsrc and l >= helper_below), None)
    if ml is not None:
        for name, lo, hi in phases:
            if lo <= ml <= hi: cur = name; break
    a = agg[cur]
    a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    a["samples"] += int(r[ix["# Samples"]] or 0)
    for s in stalls: a[s] += int(r[ix[s]] or 0)
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print(f"{'phase':14s} {'inst%':>6s} {'time%':>6s} {'smp/kinst':>9s}  top states (share of the phase's samples)")
for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    if not a["samples"]: continue
    top = sorted(((s[6:], a[s]) for s in stalls), key=lambda kv: -kv[1])[:6]
    print(f"{ph:14s} {100*a['inst']/toti:6.1f} {100*a['samples']/tot:6.1f} {1000*a['samples']/max(a['inst'],1):9.2f}  " +
          " ".join(f"{n}:{100*v/a['samples']:.0f}%" for n, v in top))

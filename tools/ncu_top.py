#!/usr/bin/env python3
"""Top SASS instructions of an ncu source page by stall samples, with the dominant stall reasons and
the CUDA source line (nvdisasm line info).  Usage: ncu_top.py src.csv kernel.cubin kernel_name [topN]"""
import csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}; cur = None; infn = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln): infn = kname in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
agg = {}
for h in st: agg[h] = sum(int(r[ix[h]] or 0) for r in data)
print({k: round(100.0 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
for r in data[:topn]:
    off = int(r[0], 16) - base
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in st), reverse=True)[:3]
    print(f"{100.0*s/tot:5.2f}%  {off:6x}  {r[1].strip()[:58]:58s} {line_of.get(off)}  " + " ".join(f"{n}:{c}" for c, n in reasons if c))

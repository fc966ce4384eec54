#!/usr/bin/env python3
"""rt-weighted instruction cost of a kernel from an ncu source page: sum over SASS instructions of
(executions x reciprocal throughput of the opcode class, measured by tools/pipebench.cu), grouped by the CUDA
source line ranges given as PHASES="name:lo-hi,..." (lines of the kernel's main source file).
Usage: ncu_cost.py src.csv kernel.cubin kernel_name main_source.cu n_units"""
import csv, os, re, subprocess, sys
from collections import defaultdict
src_csv, cubin, kname, mainsrc, units = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
RT = {"FFMA": 1, "FADD": 1, "FMUL": 1, "FFMA2": 2, "FADD2": 2, "FMUL2": 2, "SHFL": 4, "POPC": 8, "I2F": 8, "F2F": 8, "F2I": 8, "I2FP": 2,
      "MUFU": 8, "DADD": 4, "DMUL": 4, "DFMA": 4, "DSETP": 4, "BREV": 8, "FLO": 8, "REDUX": 4, "LDS": 2, "STS": 2, "LDG": 2, "STG": 2,
      "ATOMS": 4, "BAR": 2, "SYNCS": 2, "S2UR": 4, "S2R": 4}
def rt(op): return RT.get(op, 2)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}; stack = None; infn = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln): infn = kname in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = m.group(1).split("/")[-1], int(m.group(2)), m.group(3)
        # inlined: "inlined at "file", line N" chains -> take the outermost frame in the main source
        frames = [(f, l)] + [(a.split("/")[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
        stack = frames
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
cost = defaultdict(float); cnt = defaultdict(float); opc = defaultdict(lambda: defaultdict(float))
helper_below = int(os.environ.get("HELPER_BELOW", "0"))      # main-source lines below this are inlined helpers
cur = "other"
for r in sorted(data, key=lambda r: int(r[0], 16)):
    off = int(r[0], 16) - base
    n = int(r[ix["Instructions Executed"]] or 0)
    toks = r[1].strip().split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    frames = line_of.get(off) or []
    ml = next((l for f, l in frames if f == mainsrc and l >= helper_below), None)
    if ml is not None:                       # helpers / intrinsics inherit the phase of the code around them
        for name, lo, hi in phases:
            if lo <= ml <= hi: cur = name; break
    ph = cur
    cost[ph] += n * rt(op); cnt[ph] += n; opc[ph][op] += n * rt(op)
tc = sum(cost.values()); tn = sum(cnt.values())
print(f"total: {tn/units:8.0f} warp-inst/unit   {tc/units:8.0f} rt-cycles/unit (= {tc/units/4:.0f} cycles per unit per SM if the 4 sub-partitions were perfectly busy)")
for ph in sorted(cost, key=lambda k: -cost[k]):
    top = sorted(opc[ph].items(), key=lambda kv: -kv[1])[:7]
    print(f"  {ph:12s} {cnt[ph]/units:7.0f} inst  {cost[ph]/units:7.0f} rt-cyc ({100*cost[ph]/tc:4.1f}%)  " + " ".join(f"{o}:{c/units:.0f}" for o, c in top))

#!/bin/bash
# tuning: the tc16 filter under its debug modes, kernel times from an ncu launch list (KNN_QUERIES x KNN_TRAIN)
export KNN_QUERIES=${KNN_QUERIES:-262144}
for mode in ${MODES:-0 2}; do
  DSP_TC16_DEBUG=$mode ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_mode$mode.csv python tools/knn_bench.py > gpurun_out/knn_mode$mode.log 2>&1
  echo "mode $mode"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/knn_mode$mode.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg={}
for r in rows[hdr+1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e6 if u=="ns" else (v/1e3 if u=="us" else v)
    agg.setdefault(r[ki][:60],[]).append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]: print(f"  {sum(v):9.3f} ms  x{len(v):3d}  {k}")
PY
done

#!/bin/bash
# one GPU round trip: the parts of the suite named on the command line (each under its own timeout), logs to gpurun_out/
mkdir -p gpurun_out
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 300 python -m pytest "$t" -x -q > gpurun_out/chk_$name.log 2>&1
  echo "== $t rc=$? $(tail -1 gpurun_out/chk_$name.log)"
done

#!/bin/bash
# tuning: tools/quick_pipe.py under several builds of the library (tools/variants/*.so), with and without phase counters
N=${N:-100000}
for lib in "$@"; do
  echo "=== $lib"
  if [ "$lib" = "tree" ]; then unset DSP_LIB_PATH; else export DSP_LIB_PATH=$PWD/tools/variants/$lib.so; fi
  timeout 120 python tools/quick_pipe.py $N 2>&1 | grep -v "^\[prof\|^\[W"
  DSP_PROF=1 QUICK_LAYOUTS=${PROF_LAYOUT:-aligned} timeout 120 python tools/quick_pipe.py 20000 2>&1 | grep "prof pipe" | awk "NR==5"
done

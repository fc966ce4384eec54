#!/bin/bash
# one GPU round trip: ncu --set full captures of the hot kernels (each command has exited 0 without ncu before)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"frontend_pipe_kernel|knn_tc16_filter" -c 2 \
    -f -o gpurun_out/r02_bench python bench.py --utts 20000 --train-utts 20000 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-knn > gpurun_out/ncu_a.log 2>&1
ncu -i gpurun_out/r02_bench.ncu-rep --page raw --csv > gpurun_out/r02_bench_raw.csv
ncu -i gpurun_out/r02_bench.ncu-rep --page source --csv -k regex:frontend_pipe_kernel > gpurun_out/r02_pipe_src.csv
ncu --set full --clock-control none --import-source on -k regex:frontend_pipe_kernel -s 3 -c 1 \
    -f -o gpurun_out/r02_pipe_1102 python tools/config_sweep.py 20000 1102/441 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/r02_pipe_1102.ncu-rep --page raw --csv > gpurun_out/r02_pipe_1102_raw.csv
ncu -i gpurun_out/r02_pipe_1102.ncu-rep --page source --csv > gpurun_out/r02_pipe_1102_src.csv
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
rm -f gpurun_out/*.ncu-rep

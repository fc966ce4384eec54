#!/bin/bash
# Round-2 evidence in one GPU round trip (every program has exited 0 without ncu before): the bench line, the launch
# list of its timed region, ncu --set full of the hot kernels (raw + source pages), the config-4 sweep.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --no-e2e --no-cpu-baseline --no-knn > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"frontend_pipe_kernel|knn_tc16_filter" -c 2 \
    -f -o gpurun_out/r02_bench python bench.py --utts 20000 --train-utts 20000 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-knn > gpurun_out/ncu_a.log 2>&1
ncu -i gpurun_out/r02_bench.ncu-rep --page raw --csv > gpurun_out/r02_bench_raw.csv
ncu -i gpurun_out/r02_bench.ncu-rep --page source --csv -k regex:frontend_pipe_kernel > gpurun_out/r02_pipe_src.csv
ncu --set full --clock-control none --import-source on -k regex:frontend_pipe_kernel -s 3 -c 1 \
    -f -o gpurun_out/r02_pipe_1102 python tools/config_sweep.py 20000 1102/441 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/r02_pipe_1102.ncu-rep --page raw --csv > gpurun_out/r02_pipe_1102_raw.csv
ncu -i gpurun_out/r02_pipe_1102.ncu-rep --page source --csv > gpurun_out/r02_pipe_1102_src.csv
KNN_QUERIES=262144 ncu --set full --clock-control none -k regex:knn_tc16_filter -s 1 -c 1 \
    -f -o gpurun_out/r02_knn_tc16 python tools/knn_bench.py > gpurun_out/ncu_c.log 2>&1
ncu -i gpurun_out/r02_knn_tc16.ncu-rep --page raw --csv > gpurun_out/r02_knn_tc16_raw.csv
DSP_KNN_NO_TC16=1 KNN_QUERIES=262144 ncu --set full --clock-control none -k regex:knn_scan_pair -s 1 -c 1 \
    -f -o gpurun_out/r02_knn_pair python tools/knn_bench.py > gpurun_out/ncu_d.log 2>&1
ncu -i gpurun_out/r02_knn_pair.ncu-rep --page raw --csv > gpurun_out/r02_knn_pair_raw.csv
rm -f gpurun_out/*.ncu-rep
python tools/config_sweep.py 20000 256/128,1102/441,64/32,128/64,512/256,1024/512,2048/1024,352/441,441/441,529/441,661/441,793/441,882/441,1323/441,1543/441,1764/441,1984/441,2205/441,1102/132,1102/220,1102/308,1102/352,1102/529,1102/661,1102/793,1102/882,1102/1102,1102/1323 > gpurun_out/r02_config_sweep.txt 2>&1
tail -3 gpurun_out/r02_config_sweep.txt

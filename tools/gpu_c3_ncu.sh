#!/bin/bash
# ncu --set full of the general solver's kernels in a config-3 batch: per fit 8 k_pass5 / k_update5 launches
# (5 coarse, 3 full), so launch 8 is the first coarse iteration of the second fit and 13 its first full one
mkdir -p gpurun_out
run() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f \
  python tools/bench_c3.py 512 11011 0.99 > gpurun_out/$3.log 2>&1; }
run k_update5 8 r02_k_update5_coarse
run k_update5 13 r02_k_update5_full
run k_pass5 8 r02_k_pass5_coarse
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# config 5 (ppalign iteration) timing + ncu --set full of k_guess
mkdir -p gpurun_out
timeout 600 python tools/bench_c5.py 2000 > gpurun_out/r02_bench_config5.json 2> gpurun_out/r02_c5.err
tail -c 1500 gpurun_out/r02_bench_config5.json; tail -3 gpurun_out/r02_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_guess -s 2 -c 1 -o gpurun_out/r02_k_guess -f \
  python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1 > gpurun_out/r02_ncu_guess.log 2>&1
ls -la gpurun_out/r02_k_guess.ncu-rep

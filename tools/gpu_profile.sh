#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench + --set full of the two main kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1 > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spectra16 -s 2 -c 1 -o gpurun_out/r02_k_spectra16 -f \
  python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1 > gpurun_out/r02_ncu_spectra16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass2 -s 2 -c 1 -o gpurun_out/r02_k_pass2 -f \
  python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1 > gpurun_out/r02_ncu_pass2.log 2>&1
ls -la gpurun_out/r02_k_*.ncu-rep gpurun_out/r02_launches.csv

#!/bin/bash
# ncu evidence with the model's harmonic cut-off: launch list + --set full of k_spectra16 and k_pass2 (default bench shape)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_spectra16 -s 2 -c 1 -o gpurun_out/r02_k_spectra16 -f $B > gpurun_out/r02_ncu_spectra16.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_pass2 -s 2 -c 1 -o gpurun_out/r02_k_pass2 -f $B > gpurun_out/r02_ncu_pass2.log 2>&1
for f in r02_k_spectra16 r02_k_pass2; do python tools/summarize_ncu.py kernel gpurun_out/$f.ncu-rep gpurun_out/$f.md > /dev/null 2>&1; rm -f gpurun_out/$f.ncu-rep; done
ls -la gpurun_out/*.md gpurun_out/r02_launches.csv

#!/bin/bash
# round-2 ncu evidence after the solver changes: launch list of a short default bench, --set full of k_pass2 (cp.async
# ring), of a full-resolution k_pass5 / k_update5 (config 3 without the coarse levels: every launch is a full one) and
# of one coarse k_pass5; the reports are summarised on the box (tools/summarize_ncu.py) and dropped: four of them exceed
# what gpurun brings back
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --nsub 1000 --no-cpu --no-extras --e2e-nsub 64 --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_pass2 -s 2 -c 1 -o gpurun_out/r02_k_pass2 -f $B > gpurun_out/r02_ncu_pass2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_pass5 -s 8 -c 1 -o gpurun_out/r02_k_pass5_full -f python tools/bench_c3.py 512 11011 0 > gpurun_out/r02_ncu_pass5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_update5 -s 8 -c 1 -o gpurun_out/r02_k_update5 -f python tools/bench_c3.py 512 11011 0 > gpurun_out/r02_ncu_update5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_pass5 -s 16 -c 1 -o gpurun_out/r02_k_pass5_coarse -f python tools/bench_c3.py 512 11011 0.99 > gpurun_out/r02_ncu_pass5c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pass5|k_update5|k_spectra|k_guess|k_coarse|k_model_info|k_count" -s 40 -c 60 --csv --log-file gpurun_out/r02_launches_config3.csv python tools/bench_c3.py 512 11011 0.99 > gpurun_out/r02_ncu_c3l.log 2>&1
for f in r02_k_pass2 r02_k_pass5_full r02_k_update5 r02_k_pass5_coarse; do python tools/summarize_ncu.py kernel gpurun_out/$f.ncu-rep gpurun_out/$f.md > /dev/null 2>&1; rm -f gpurun_out/$f.ncu-rep; done
ls -la gpurun_out/*.md gpurun_out/r02_launches*.csv

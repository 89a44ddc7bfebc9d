#!/bin/bash
# config 3 A/B: coarse stage off / on, both flag sets
mkdir -p gpurun_out
for fl in ${C3_FLAGS:-11011 11111}; do
  timeout 600 python tools/bench_c3.py ${C3_NSUB:-512} $fl ${C3_FRACS:-0,0.99} 2>> gpurun_out/r02_c3.err | tee -a gpurun_out/r02_bench_config3.jsonl
done
tail -5 gpurun_out/r02_c3.err

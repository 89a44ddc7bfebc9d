#!/bin/bash
mkdir -p gpurun_out
PP_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-nsub 256 --e2e-steps 1 > gpurun_out/r02_trace.json 2> gpurun_out/r02_trace.err
grep -n "pp_fit_batch" gpurun_out/r02_trace.err | awk 'BEGIN{c=0} /inputs staged/{c++} {if (c==4 || c==5) print}' | head -60
tail -c 600 gpurun_out/r02_trace.json

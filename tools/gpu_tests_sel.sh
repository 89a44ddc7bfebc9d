#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest $@ -m gpu -q > gpurun_out/r02_gpu_tests_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/r02_gpu_tests_sel.log | tail -30

#!/bin/bash
# full GPU validation: parity suite + default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1
tail -5 gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 3000 gpurun_out/r02_bench.json

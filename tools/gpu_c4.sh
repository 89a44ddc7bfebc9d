#!/bin/bash
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then timeout 900 python tools/bench_c4.py 1000000 10000 > gpurun_out/r02_bench_config4_n1.json 2> gpurun_out/r02_c4_n1.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_c4.py 1000000 10000 > gpurun_out/r02_bench_config4_n$N.json 2> gpurun_out/r02_c4_n$N.err; fi
cat gpurun_out/r02_bench_config4_n$N.json; tail -2 gpurun_out/r02_c4_n$N.err

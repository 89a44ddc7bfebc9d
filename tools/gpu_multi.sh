#!/bin/bash
# usage: gpu_multi.sh N  -- multi-GPU tests + bench on N GPUs of one box
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_gpu_multi_n$N.log 2>&1
tail -3 gpurun_out/r02_gpu_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read())
print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'i16', d['e2e_i16']['value'])
print('strong', d['strong_scaling']); print(d['config']['converged'], d['config']['dDM_pull_rms'])
PY
tail -3 gpurun_out/r02_bench_n$N.err

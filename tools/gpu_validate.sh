#!/bin/bash
# full GPU validation: parity suite + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gpu_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_gpu_tests.log | tail -25
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('facade', d.get('facade')); print('config3', d.get('config3'))
print('roofline frac', d['roofline']['frac'], 'step_frac', d['roofline']['step_frac_actual_bytes'], {k:v.get('ms') for k,v in d['roofline']['kernels'].items()})
PY
tail -3 gpurun_out/r02_bench.err

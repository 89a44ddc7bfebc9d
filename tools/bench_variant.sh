#!/bin/bash
# usage: tools/bench_variant.sh <lib.so> : run the short bench with an alternative build of the library
cp pulseportraiture_b200/libppb200.so /tmp/_orig.so
cp "$1" pulseportraiture_b200/libppb200.so 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 3 --nsub 4000 --no-cpu --e2e-nsub 512 --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('$1', round(d['value']), 'TOA/s', round(d['ms_per_step'],2), 'ms/step | spectra %.2f ms (%.0f GB/s)  pass %.2f ms (%.0f GB/s) guess %.2f update %.2f total %.2f' % (k['k_spectra']['ms'], k['k_spectra']['achieved_gbs'], k['k_pass2']['ms'], k['k_pass2']['achieved_gbs'], k['k_guess']['ms'], k['k_update2']['ms'], d['roofline']['ms_total']))"
cp /tmp/_orig.so pulseportraiture_b200/libppb200.so

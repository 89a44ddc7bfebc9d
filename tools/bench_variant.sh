#!/bin/bash
# usage: tools/bench_variant.sh <lib.so> : run the short bench with an alternative build of the library
cp pulseportraiture_b200/libppb200.so /tmp/_orig.so
cp "$1" pulseportraiture_b200/libppb200.so
timeout 300 python bench.py --steps 3 --warmup 3 --nsub 4000 --no-cpu --e2e-nsub 512 --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', round(d['value']), round(d['ms_per_step'],2), {k:round(r[k],2) for k in ('achieved','ms_pass','ms_spectra','ms_guess','ms_update','ms_total')})"
cp /tmp/_orig.so pulseportraiture_b200/libppb200.so

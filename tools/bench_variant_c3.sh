#!/bin/bash
# usage: tools/bench_variant_c3.sh <lib.so> : config-3 bench with an alternative build of the library
cp pulseportraiture_b200/libppb200.so /tmp/_orig.so
cp "$1" pulseportraiture_b200/libppb200.so 2>/dev/null
timeout 300 python tools/bench_c3.py 512 11011 2>&1 | tail -1 | cut -c60-330
cp /tmp/_orig.so pulseportraiture_b200/libppb200.so

#!/bin/bash
# usage: tools/variant_probe.sh <nbins> lib1.so lib2.so ... : tools/anynbin_probe.py with alternative builds of the library
NB=$1; shift
cp pulseportraiture_b200/libppb200.so /tmp/_orig.so
for v in orig "$@"; do
  [ "$v" != orig ] && cp "$v" pulseportraiture_b200/libppb200.so
  timeout 300 python tools/anynbin_probe.py 1000 $NB 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$v', ' | '.join('nbin %d: %.0f TOA/s spectra %.3f ms' % (c['nbin'], c['TOAs_per_s'], c['ms_spectra']) for c in d['cases']))"
done
cp /tmp/_orig.so pulseportraiture_b200/libppb200.so

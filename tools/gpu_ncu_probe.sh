#!/bin/bash
# usage: gpu_ncu_probe.sh <launch-skip> <name>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectra -s $1 -c 1 -o gpurun_out/$2 -f ./build_probe/spectra_probe 200 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log

#!/bin/bash
# GPU call 1 (round 2): k_spectra variants side by side + ncu of the default and of the 8-CTA variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_call1_smi.txt
./build_probe/spectra_probe 1000 > gpurun_out/r02_probe1.txt 2>&1
cat gpurun_out/r02_probe1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectra -s 1 -c 1 -o gpurun_out/r02_spectra_base -f ./build_probe/spectra_probe 200 > gpurun_out/r02_ncu_base.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectra -s 25 -c 1 -o gpurun_out/r02_spectra_v8 -f ./build_probe/spectra_probe 200 > gpurun_out/r02_ncu_v8.log 2>&1
tail -3 gpurun_out/r02_ncu_v8.log

#!/bin/bash
mkdir -p gpurun_out
timeout 90 ./build_probe/spectra_probe ${1:-1000} 2>&1 | tee gpurun_out/r02_probe_last.txt

#!/bin/bash
mkdir -p gpurun_out
./build_probe/cvt_probe > gpurun_out/r02_cvt_probe.txt 2>&1
cat gpurun_out/r02_cvt_probe.txt
./build_probe/spectra_probe 1000 > gpurun_out/r02_probe2.txt 2>&1
cat gpurun_out/r02_probe2.txt

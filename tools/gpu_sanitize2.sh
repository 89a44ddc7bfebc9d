#!/bin/bash
# compute-sanitizer over the tests that touch the kernels changed last in round 2: k_pass2 / k_pass5 with the cp.async
# ring, the coarse levels (harmonic cut-off, channel stride), k_update5's scalar form, k_model_info, staged float64 rows
mkdir -p gpurun_out
SEL="tests/test_gpu_round2.py::test_general_solver_coarse_stage_same_optimum tests/test_gpu_round2.py::test_float64_input_needs_no_host_pass_and_is_bit_identical tests/test_gpu_golden_v2.py::test_every_flag_pattern_sigma_1p5_against_reference tests/test_gpu_parity.py::test_every_supported_nbin tests/test_gpu_parity.py::test_masks_errs_dmguess_nufit_modes tests/test_gpu_bounds.py"
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -m gpu -q -x > gpurun_out/r02_san2_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_san2_$tool.log | tail -3
done

#!/bin/bash
# compute-sanitizer over the tests that touch the round-2 kernels (k_spectra16, Bluestein rows, FROMSPEC emit,
# float64 conversion, saddle-free solver step, response multiply).  Logs -> gpurun_out/r02_san_*.log
mkdir -p gpurun_out
SEL="tests/test_gpu_anynbin.py::test_any_nbin_batch_with_guess_int16_float64_and_align tests/test_gpu_round2.py::test_float64_input_needs_no_host_pass_and_is_bit_identical tests/test_gpu_golden_v2.py::test_instrumental_response_against_reference tests/test_gpu_parity.py::test_int16_input_matches_decoded_float32 tests/test_gpu_parity.py::test_nonfinite_and_empty_subints_do_not_poison_the_batch tests/test_gpu_golden_v2.py::test_every_flag_pattern_sigma_1p5_against_reference tests/test_gpu_align.py::test_fused_fit_and_align_matches_two_steps"
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -m gpu -q -x > gpurun_out/r02_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_san_$tool.log | tail -3
done
# k_spectra16 at the benchmark shape (2048-bin rows, masks, int16, align) under racecheck
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest "tests/test_gpu_parity.py::test_every_supported_nbin" "tests/test_gpu_parity.py::test_masks_errs_dmguess_nufit_modes" -m gpu -q > gpurun_out/r02_san_racecheck_nbin.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_san_racecheck_nbin.log | tail -2

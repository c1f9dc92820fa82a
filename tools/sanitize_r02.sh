#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
# round 2: the fused ADE kernel (K1-ADE), K1's box variants beside the ADE list kernels, the cut-first schedule, the
# checkpoint path, K5 after its changes, and a slice of the random configurations
SEL4='(test_fused_ade_kernel_matches_oracle and (auto or r1_flat_chunk3 or r2_strips_graph)) or (test_ade_layouts_match_oracle and (ade_two or ade_dense)) or test_fused_ade_survives or (test_checkpoint_and_resume and ade_dense) or (test_resident_kernel_matches_oracle and (odd_geometry or nonuniform) and 0) or (test_random_configuration_matches_oracle and (0] or 1] or 3] or 7]))'
for tool in memcheck initcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 --target-processes all \
      python -m pytest tests/test_parity_gpu.py tests/test_fuzz_parity_gpu.py -m gpu -x -q -k "$SEL4" > gpurun_out/sanitizer_r02_$tool.log 2>&1
  echo "r02-$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_r02_$tool.log | tr '\n' ' ')"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "cut_planes_first or velocity_membrane" > gpurun_out/sanitizer_r02_memcheck_multi.log 2>&1
echo "r02-memcheck-multi exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_r02_memcheck_multi.log | tr '\n' ' ')"
# later in round 2: Mur / radiation planes inside K6's tiles, 24 point sources inside K5 / K6 / K1, decomposition fuzz
SEL5='test_pipelined_kernel_applies_mur or (test_pipelined_kernel_picks_tiles and (shape0 or shape2)) or test_chunk_kernels_take_a_phased_array or (test_random_configuration_is_decomposition_invariant and (102] or 107] or 119]))'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 --target-processes all \
      python -m pytest tests/test_parity_gpu.py tests/test_fuzz_parity_gpu.py -m gpu -x -q -k "$SEL5" > gpurun_out/sanitizer_r02b_$tool.log 2>&1
  echo "r02b-$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_r02b_$tool.log | tr '\n' ' ')"
done

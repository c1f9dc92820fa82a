#!/usr/bin/env bash
# compute-sanitizer pass over a few small parity cases (memcheck + racecheck + initcheck); writes gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
SEL='test_matches_oracle_and_golden and (odd_geometry_pml or ade_two or directional or pml_radiation) and (march_r2_separate_k3 or march_r1_chunk5_fused or naive)'
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "peer_store and block_pml" > gpurun_out/sanitizer_memcheck_p2p.log 2>&1
echo "memcheck-p2p exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_memcheck_p2p.log | tr '\n' ' ')"
# the chunk kernels (shared-memory-resident K5, step-pipelined K6) and ADE across slab cuts
SEL2='(test_resident_kernel_matches_oracle and (odd_geometry or two_sponges or nonuniform) and 0) or (test_pipelined_kernel_matches_oracle and (odd_geometry or nonuniform) and shape_opts1) or (test_ade_materials_across_slab_cuts and copy and 2)'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --target-processes all \
      python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py -m gpu -x -q -k "$SEL2" > gpurun_out/sanitizer_chunk_$tool.log 2>&1
  echo "chunk-$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_chunk_$tool.log | tr '\n' ' ')"
done
SEL3='(test_matches_oracle_and_golden and (march_r1_flat or march_r2_flat_fused) and (odd_geometry or ade_dense or nonuniform_block)) or (test_ade_layouts_match_oracle and ade_dense_layers) or (test_pipelined_kernel_matches_oracle and shape_opts3 and (odd_geometry or two_sponges))'
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL3" > gpurun_out/sanitizer_flat_$tool.log 2>&1
  echo "flat-$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_flat_$tool.log | tr '\n' ' ')"
done
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

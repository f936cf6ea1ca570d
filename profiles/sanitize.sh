#!/bin/bash
# compute-sanitizer evidence: memcheck and racecheck of the BASELINE config-1 pipeline
# (reorder -> candidates -> thresholds -> fp32 enumeration with early exit, tail splitting and top-k lists,
# fp64 re-score), of the small parity tests and of the pair path (context tables, pair kernel, flat list of
# irregular trees at K = 8, programs-only build), through the C ABI.  Run on the GPU box:
#   bash profiles/sanitize.sh          -> gpurun_out/sanitize_{memcheck,racecheck}.log
# The in-place convolution (read-before-overwrite across half-warps, __syncwarp) and the tail splitting
# (atomics + __threadfence) are what racecheck is pointed at.
set -u
mkdir -p gpurun_out
SEL="tests/test_gpu_search.py::test_config1_pipeline_fp32_equals_check_mode tests/test_gpu_parity.py::test_topk_matches_oracle_ranking tests/test_gpu_parity.py::test_tree_range_sharding_equals_single_range tests/test_gpu_pairs.py::test_topk_pairs_equals_programs_and_check_mode tests/test_gpu_pairs.py::test_all_trees_pairs_vs_oracle_and_programs"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest $SEL -x -q -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit code: $?" >> gpurun_out/sanitize_$tool.log
  tail -n 6 gpurun_out/sanitize_$tool.log
done

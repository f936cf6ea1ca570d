#!/bin/bash
# Substitute for compute-sanitizer memcheck (closed on this GPU pool): the GPU test suite on the bounds-checked build
# (make -C bamse_b200/csrc DEBUG=1 -> libbamse_cuda_dbg.so: every table-row / list index computed in a kernel is
# checked against its extent and a violation traps).  Run on the GPU box: bash profiles/bounds_check.sh
set -u
mkdir -p gpurun_out
[ -f bamse_b200/libbamse_cuda_dbg.so ] || make -C bamse_b200/csrc DEBUG=1 > /dev/null
BAMSE_CUDA_LIB=$PWD/bamse_b200/libbamse_cuda_dbg.so python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/bounds_check.log 2>&1
echo "exit code: $?" >> gpurun_out/bounds_check.log
tail -n 4 gpurun_out/bounds_check.log

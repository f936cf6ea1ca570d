#!/bin/bash
# round-2 profiling pass (one gpurun call): plain run first, then the ncu launch list of the same command,
# then one `--set full` capture of the pair enumeration kernel and of the table-building kernels.
set -u
mkdir -p gpurun_out
CFG=${1:-c2}
CMD="python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --lean"
$CMD > gpurun_out/prof_plain_$CFG.json 2> gpurun_out/prof_plain_$CFG.err || { echo "plain run failed"; tail gpurun_out/prof_plain_$CFG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_$CFG.csv $CMD > gpurun_out/prof_ncu1_$CFG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_enumerate_pairs -s 4 -c 1 -f -o gpurun_out/r02_pairs_$CFG $CMD > gpurun_out/prof_ncu2_$CFG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_forest_level|k_ctx_level|k_tables_fast' -s 40 -c 10 -f -o gpurun_out/r02_build_$CFG $CMD > gpurun_out/prof_ncu3_$CFG.log 2>&1
ls -la gpurun_out/*.ncu-rep

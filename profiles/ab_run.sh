#!/bin/bash
# A/B of kernel variants on one B200: profiles/ab_run.sh <tag> <config> "<lib or ->:<PROGTAB>" ...
# writes gpurun_out/ab_<tag>_<i>.json (bench lines, no cpu_baseline leg)
tag=$1; cfg=$2; shift 2
i=0
for v in "$@"; do
    lib=${v%%:*}; pt=${v##*:}
    if [ "$lib" = "-" ]; then unset BAMSE_CUDA_LIB; else export BAMSE_CUDA_LIB=$PWD/variants/$lib/libbamse_cuda.so; fi
    BAMSE_PROGTAB=$pt python bench.py --config $cfg --no-cpu-baseline --steps ${STEPS:-20} --warmup 3 > gpurun_out/ab_${tag}_$i.json 2> gpurun_out/ab_${tag}_$i.err
    echo "$v $(python -c "import json,sys; d=json.loads(open('gpurun_out/ab_${tag}_$i.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'].get('executed'))" 2>&1 | tail -1)"
    i=$((i+1))
done

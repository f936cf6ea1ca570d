#!/bin/bash
# scaling lines of one build: bash profiles/scale_run.sh <N> <configs...>   (under gpurun --gpus N)
# writes gpurun_out/r02_scale_<config>_<N>.json (last line = the bench line)
set -u
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  EXTRA="--steps 5 --warmup 3 --no-cpu-baseline"
  [ "$c" = "c4" ] && EXTRA="--steps 1 --warmup 1 --lean --no-cpu-baseline"
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --config $c $EXTRA > gpurun_out/r02_scale_${c}_${N}.json 2> gpurun_out/r02_scale_${c}_${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c $EXTRA > gpurun_out/r02_scale_${c}_${N}.json 2> gpurun_out/r02_scale_${c}_${N}.err
  fi
  grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/r02_scale_${c}_${N}.err | tail -n 5
  python - "$c" "$N" <<'PY'
import json, sys
c, n = sys.argv[1], sys.argv[2]
lines = [l for l in open("gpurun_out/r02_scale_%s_%s.json" % (c, n)) if l.startswith("{")]
if lines:
    d = json.loads(lines[-1]); ex = d["roofline"]["executed"]
    print("%s N=%s value %.1fM (%.3f ms/step) e2e %.1fM (%.3f ms) tab %.3f enum %.3f ms" % (c, n, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], ex["table_build_ms"], ex["enumeration_ms"]))
PY
done

#!/bin/bash
# same-box comparison of several environment configurations on the replayed config-2 step:
#     ab_cfgs.sh <tag> <reps> "<VAR=v VAR2=v ...>" "<...>" ...      ("-" = no switch)
TAG=$1; REPS=$2; shift 2
for r in $(seq 1 $REPS); do i=0; for cfg in "$@"; do i=$((i+1))
  [ "$cfg" = "-" ] && cfg=""
  env $cfg python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_c${i}_$r.json 2>gpurun_out/${TAG}_c${i}_$r.err
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_c${i}_$r.json').read().strip().splitlines()[-1]);print('[$cfg] run $r: %.2f samples/s  %.3f ms/step  gemm %.2f ms' % (d['value'],d['ms_per_step'],d['roofline']['gemm_ms_per_step']))" || tail -3 gpurun_out/${TAG}_c${i}_$r.err
done; done

#!/bin/bash
# same-box A/B of one environment switch on the replayed config-2 step:  ab_env.sh <tag> <VAR> <valueA> <valueB> [reps]
TAG=$1; VAR=$2; A=$3; B=$4; REPS=${5:-2}
for r in $(seq 1 $REPS); do
  env $VAR=$A python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_${VAR}_${A}_$r.json 2>gpurun_out/${TAG}_${VAR}_${A}_$r.err
  env $VAR=$B python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_${VAR}_${B}_$r.json 2>gpurun_out/${TAG}_${VAR}_${B}_$r.err
done
for v in $A $B; do for r in $(seq 1 $REPS); do
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_${VAR}_${v}_$r.json').read().strip().splitlines()[-1]);print('$VAR=$v run $r: %.2f samples/s  %.3f ms/step  gemm %.0f TFLOP/s %.2f ms' % (d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step']))" || tail -3 gpurun_out/${TAG}_${VAR}_${v}_$r.err
done; done

#!/bin/bash
# same-box A/B of two source trees (each with its own built library) on the replayed config-2 step: ab_tree.sh <tag> <treeA> <treeB> [reps]
TAG=$1; A=$2; B=$3; REPS=${4:-2}
OUT=$PWD/gpurun_out
for r in $(seq 1 $REPS); do
  for v in A B; do
    if [ $v = A ]; then T=$A; else T=$B; fi
    (cd $T && python bench.py --no-extras --no-cpu-baseline > $OUT/${TAG}_$v$r.json 2> $OUT/${TAG}_$v$r.err)
    python -c "import json;d=json.loads(open('$OUT/${TAG}_$v$r.json').read().strip().splitlines()[-1]);print('$v ($T) run $r: %.2f samples/s  %.3f ms/step  launches/step %.0f' % (d['value'],d['ms_per_step'],d['gpu_launches_per_step']))" || tail -3 $OUT/${TAG}_$v$r.err
  done
done

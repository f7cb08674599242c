#!/bin/bash
# same-box A/B of two builds of the library on the replayed config-2 step:  ab_lib.sh <tag> <libA.so> <libB.so> [reps]
TAG=$1; A=$2; B=$3; REPS=${4:-2}
LIB=bubbleformer_b200/libbubbleformer_b200.so
cp $LIB /tmp/lib_keep.so
for r in $(seq 1 $REPS); do
  for v in A B; do
    if [ $v = A ]; then cp $A $LIB; else cp $B $LIB; fi
    python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_$v$r.json 2> gpurun_out/${TAG}_$v$r.err
    python -c "import json;d=json.loads(open('gpurun_out/${TAG}_$v$r.json').read().strip().splitlines()[-1]);print('$v run $r: %.2f samples/s  %.3f ms/step' % (d['value'],d['ms_per_step']))" || tail -3 gpurun_out/${TAG}_$v$r.err
  done
done
cp /tmp/lib_keep.so $LIB

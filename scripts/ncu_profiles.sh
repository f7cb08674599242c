#!/bin/bash
# Round profile captures (run under gpurun, one GPU).  Outputs under gpurun_out/; summaries are copied to profiles/
# by scripts/ncu_summarise.py.  Launch indices: 3 eager warm-up steps + 3 warm-up replays precede the timed region,
# 234 GEMM launches and ~1100 launches of all kernels per step.
set -x
TAG=${1:-r1q}
CMD="python bench.py --no-cpu-baseline --steps 2 --warmup 3"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
# (1) every launch of one replayed step with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6600 --launch-count 1300 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# (2) DRAM traffic of every GEMM launch of one replayed step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_tcgen05 --launch-skip 1404 --launch-count 234 --csv --log-file gpurun_out/gemm_traffic_$TAG.csv \
    $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
# (3) full sections for the first block GEMMs of that step (QKV+LN, out-projection, QKV+LN, out-projection, fc1, fc2)
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 1407 --launch-count 6 \
    -o gpurun_out/prof_gemm_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log

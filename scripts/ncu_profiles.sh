#!/bin/bash
# Round profile captures (run under gpurun, one GPU).  Outputs under gpurun_out/; scripts/ncu_summarise.py turns them
# into the committed summaries under profiles/.  `bench.py --profile --steps 2` runs ONE eager warm-up step (~1100
# launches, 234 of them GEMMs), captures the training-step graph and replays it twice (734 launches, 234 GEMMs each).
set -x
TAG=${1:-r1q}
CMD="python bench.py --profile --steps 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
# (1) every launch with its device time (cold cache, serialised: compare shares); the summary keeps one replayed step
ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 2800 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# (2) DRAM traffic of every GEMM launch of the first replayed step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_tcgen05 --launch-skip 234 --launch-count 234 --csv --log-file gpurun_out/gemm_traffic_$TAG.csv \
    $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
# (3) full sections for the first block GEMMs of that step (QKV+LN, out-projection, QKV+LN, out-projection, fc1, fc2)
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 237 --launch-count 6 \
    -o gpurun_out/prof_gemm_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log

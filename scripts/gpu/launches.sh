#!/bin/bash
# ncu launch list (device time + DRAM bytes per launch) of one replayed config-2 training step:  launches.sh <tag>
TAG=$1
CMD="python bench.py --profile --steps 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-count 3200 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log; ls -la gpurun_out/launches_$TAG.csv

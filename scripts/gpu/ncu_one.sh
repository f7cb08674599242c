#!/bin/bash
# --set full capture (with source) of ONE micro.py kernel: ncu_one.sh <tag> <micro name> <kernel regex>
TAG=$1; NAME=$2; RE=$3
ncu --set full --clock-control none --import-source on -k regex:"$RE" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_${NAME}_$TAG -f python scripts/micro.py $NAME --iters 1 > gpurun_out/${TAG}_ncu_$NAME.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_$NAME.log; ls -la gpurun_out/prof_${NAME}_$TAG.ncu-rep

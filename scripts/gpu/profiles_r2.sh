#!/bin/bash
# Round-2 profile captures (one GPU, under gpurun).  `bench.py --profile --steps 2`: one eager warm-up step, the captured
# training-step graph, two replays.  Outputs under gpurun_out/; scripts/ncu_summarise_r2.py writes the committed summaries.
set -x
TAG=${1:-r2}
CMD="python bench.py --profile --steps 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
# (1) every launch of the run: device time + DRAM bytes (cold cache, serialised -> compare shares; bytes are per launch)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-count 3200 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# (2) full sections + source for the dominant kernels of the replayed step: QKV+LN GEMM, out-projection (residual) GEMM,
#     fc1 GEMM, attention backward (L = 32), norm1 backward apply, norm1 forward apply
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attn_fast_bwd|inorm_bwd_apply|inorm_apply" \
    --launch-skip 1450 --launch-count 40 -o gpurun_out/prof_top_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/prof_top_$TAG.ncu-rep

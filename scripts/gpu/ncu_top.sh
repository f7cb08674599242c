#!/bin/bash
# --set full captures (with source) of the dominant kernels at config-2 shapes, one launch each after two warm-ups
TAG=${1:-r2h}
NAMES="gemm_qkv_ln gemm_fc1d gemm_dmul gemm_resid_stats attn_bwd attn_fwd inorm_bwd2_add inorm_apply"
python scripts/micro.py $NAMES --iters 3 > gpurun_out/${TAG}_micro.txt 2>&1; cat gpurun_out/${TAG}_micro.txt
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attn_fast|inorm_bwd_apply|inorm_apply" \
    --launch-count 24 -o gpurun_out/prof_top_$TAG -f python scripts/micro.py $NAMES --iters 1 > gpurun_out/${TAG}_ncu_top.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_top.log; ls -la gpurun_out/prof_top_$TAG.ncu-rep

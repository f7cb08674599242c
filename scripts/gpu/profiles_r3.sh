#!/bin/bash
# Final round-2 evidence on one GPU: bench line, CUDA-event step breakdown, ncu launch list of one replayed step (time +
# DRAM bytes per launch), --set full captures (with source) of the dominant kernels at config-2 shapes, each summarised on
# the box (scripts/ncu_kernel_summary.py) so that only text travels back (gpurun_out/ is capped at 64 MiB).
#   profiles_r3.sh <tag>      -> gpurun_out/;  then here: python scripts/ncu_summarise_r2.py <tag>
TAG=${1:-r3}
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python scripts/profile_step.py > gpurun_out/${TAG}_step_breakdown.txt 2>&1; echo "breakdown rc=$?"
bash scripts/gpu/launches.sh $TAG
cap() {   # cap <micro name> <kernel regex>
  bash scripts/gpu/ncu_one.sh $TAG $1 $2 > /dev/null 2>&1
  python scripts/ncu_kernel_summary.py gpurun_out/prof_$1_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_$1.txt 2>&1
  rm -f gpurun_out/prof_$1_$TAG.ncu-rep
  head -4 gpurun_out/${TAG}_ncu_$1.txt
}
for k in gemm_qkv_ln gemm_fc1d gemm_resid_stats; do cap $k gemm_tcgen05; done
cap attn_bwd attn_fast_bwd
cap attn_fwd attn_fast_fwd
cap inorm_bwd2_add inorm_bwd_apply
cap inorm_apply inorm_apply
du -sh gpurun_out

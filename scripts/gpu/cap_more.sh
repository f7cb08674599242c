cap() {
  bash scripts/gpu/ncu_one.sh r3i $1 $2 > /dev/null 2>&1
  python scripts/ncu_kernel_summary.py gpurun_out/prof_$1_r3i.ncu-rep 22 > gpurun_out/r3i_ncu_$1.txt 2>&1
  rm -f gpurun_out/prof_$1_r3i.ncu-rep
}
cap patch_in patch_in_mma
cap attn_t_bwd attn_fast_bwd
cap patch_wgrad_f16 patch_wgrad_mma

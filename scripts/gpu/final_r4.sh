#!/bin/bash
# Evidence of the final build: bench line, CUDA-event step breakdown, ncu launch list (time + DRAM bytes) of one replayed step.
TAG=${1:-r4}
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python scripts/profile_step.py > gpurun_out/${TAG}_step_breakdown.txt 2>&1; echo "breakdown rc=$?"
bash scripts/gpu/launches.sh $TAG
python scripts/gemm_vs_cublas.py > gpurun_out/${TAG}_gemm_vs_cublas.txt 2>&1
du -sh gpurun_out

#!/bin/bash
# Same-box A/B of the replayed training step on an environment switch (run under gpurun, one GPU):
#     bash scripts/ab_step.sh BF_GEMM_RESIDENT 0 1        # alternates the two values twice, prints ms per step
# Micro-benchmarks keep operands in L2 and mis-rank variants of the short kernels; decisions are taken on this number
# (noise ~0.05 ms on one box, ~0.3 ms between boxes).
VAR=$1; A=$2; B=$3
for v in $A $B $A $B; do
  env $VAR=$v timeout 200 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['ms_per_step'], 3), 'ms/step')"
done

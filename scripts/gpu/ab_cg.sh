#!/bin/bash
# same-box A/B of the GEMM kernel variants (BF_GEMM_CG=1: single-CTA tiles, default: CTA pairs) on the replayed config-2 step
TAG=${1:-r2b}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "test_gemm" > gpurun_out/${TAG}_gemm.log 2>&1; echo "gemm rc=$?"; tail -25 gpurun_out/${TAG}_gemm.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "not test_gemm" > gpurun_out/${TAG}_rest.log 2>&1; echo "rest rc=$?"; tail -8 gpurun_out/${TAG}_rest.log
for r in a b; do
  BF_GEMM_CG=1 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_cg1$r.json 2>gpurun_out/${TAG}_cg1$r.err
  python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_cg2$r.json 2>gpurun_out/${TAG}_cg2$r.err
done
for f in cg1a cg2a cg1b cg2b; do
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'])"
done
tail -5 gpurun_out/${TAG}_cg2a.err

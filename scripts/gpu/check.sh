#!/bin/bash
# Quick check of a kernel change on one GPU: GPU test suite, then the config-2 bench line without the extra sections.
#   check.sh <tag> [micro names...]
TAG=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
if [ $# -gt 0 ]; then python scripts/micro.py "$@" 2>&1 | tee gpurun_out/${TAG}_micro.txt; fi
python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]);print('%.2f samples/s  %.3f ms/step  gemm %.0f TFLOP/s %.2f ms' % (d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step']))" || tail -5 gpurun_out/${TAG}_bench.err

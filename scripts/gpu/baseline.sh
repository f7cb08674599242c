#!/bin/bash
# State-of-the-build check on one GPU: full GPU test suite, the driver's bench line, per-op CUDA-event breakdown.
#   baseline.sh <tag>
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python scripts/profile_step.py > gpurun_out/${TAG}_profile_step.txt 2>&1; echo "profile rc=$?"
head -3 gpurun_out/${TAG}_profile_step.txt

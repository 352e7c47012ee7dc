#!/bin/bash
# quick GPU iteration: parity tests (stop at first failure) + phase timeline of the fused decode kernel
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/timeline.py --window ${WINDOWS:-4} --out $OUT/timeline.json > $OUT/timeline.txt 2>&1; cat $OUT/timeline.txt | tail -40

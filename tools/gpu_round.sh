#!/bin/bash
# One gpurun call: GPU parity tests, bench line, phase timeline, ncu launch list + full capture.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json
timeout 300 python tools/timeline.py --out $OUT/timeline.json > $OUT/timeline.txt 2>&1; cat $OUT/timeline.txt
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 4 -c 2 -f -o $OUT/decode_mega \
      python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
  ls -la $OUT
fi

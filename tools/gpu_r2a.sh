#!/bin/bash
# round 2, first GPU call: batched fused kernel parity, then the whole GPU suite, then a bench line
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "batch or chunk_size or out_of_range" > $OUT/pytest_batch.log 2>&1; echo "pytest batch exit $?"; tail -25 $OUT/pytest_batch.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err

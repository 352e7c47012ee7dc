#!/bin/bash
# round 2: bench lines of the three workloads on one GPU + the GPU suite
OUT=gpurun_out/${1:-r02h}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 900 python bench.py --config cfg3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "cfg3 exit $?"; cat $OUT/bench_cfg3.json; tail -3 $OUT/bench_cfg3.err
timeout 900 python bench.py --config cfg4 > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "cfg4 exit $?"; cat $OUT/bench_cfg4.json; tail -3 $OUT/bench_cfg4.err

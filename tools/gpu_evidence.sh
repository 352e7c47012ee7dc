#!/bin/bash
# round 2 evidence run (one GPU): launch list of the default bench, ncu --set full captures of every kernel family on the
# path, bench lines of the three workloads.   gpurun --timeout 2700 -- 'bash tools/gpu_evidence.sh r02c'
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
# 1. launch list of the default bench command (per-launch device times: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?"
# 2. full captures (one launch each): fused decode kernels, tcgen05 GEMM, prefill / perceiver attention, split-K epilogues
cap() {  # name regex skip extra-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
}
cap decode_mega decode_mega 4 python bench.py --steps 2 --warmup 3 --no-cpu
cap gemm_tc gemm_tc_kernel 400 python bench.py --steps 2 --warmup 3 --no-cpu
cap attention "attention_row_kernel" 40 python bench.py --steps 2 --warmup 3 --no-cpu
cap splitk_ln splitk_ln_epilogue 40 python bench.py --steps 2 --warmup 3 --no-cpu
cap decode_batch decode_batch 2 python tools/batch_bench.py --rows 8 --tokens 12
cap kv_attention kv_attention_kernel 10 python tools/kv_bench.py
ls -la $OUT
cap pc_attention_tc pc_attention_tc 2 python bench.py --steps 2 --warmup 3 --no-cpu
cap vocoder_conv1d conv1d_kernel 70 python tools/vocoder_bench.py --no-cpu --reps 2
# 3. bench lines of the three workloads (no profiler), phase timeline of the single-row kernel
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 600 python bench.py --config cfg3 --no-cpu > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 600 python bench.py --config cfg4 --no-cpu > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "bench cfg4 exit $?"
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; echo "timeline exit $?"
timeout 300 python tools/batch_bench.py --rows 8 --tokens 64 > $OUT/batch_bench.json 2>&1; echo "batch bench exit $?"
timeout 300 python tools/vocoder_bench.py > $OUT/stage_bench.json 2> $OUT/stage_bench.err; echo "stage bench exit $?"
rm -f $OUT/*.ncu-rep.tmp
ls -la $OUT

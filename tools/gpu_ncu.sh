#!/bin/bash
# one full ncu capture of the fused decode kernel (source-level stalls) + bench line
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 4 -c 1 -f -o $OUT/decode_mega \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json
ls -la $OUT

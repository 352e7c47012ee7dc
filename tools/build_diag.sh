#!/bin/bash
# Debug libraries of the fused single-row decode kernel:
#   DIAG_MODE=1|2 tools/build_diag.sh            wait diagnostics (mega_dev.cuh GV_WAIT_DIAG, tools/wait_diag.py)
#   DIAG_MODE=0 tools/build_diag.sh -DGV_PROG    per-warp progress markers, barrier counts and the barrier-slip detector
#                                                 (read by tools/hang_dump.py while a launch is stuck)
# Any instrumentation shifts the timing: a hang that needs the production schedule is best looked at with the production
# library under tools/hang_dump.py (exchange tags + arrival counters copied out on a side stream).
set -e
cd "$(dirname "$0")/.."
python -m genvc_b200.build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DGV_WAIT_DIAG=${DIAG_MODE:-1} "$@" \
     -c genvc_b200/csrc/decode_mega.cu -o genvc_b200/build/decode_mega_diag${DIAG_MODE:-1}.o
B=genvc_b200/build
nvcc -shared -o genvc_b200/libgenvc_diag${DIAG_MODE:-1}.so $B/api.o $B/ops.o $B/decode_mega_diag${DIAG_MODE:-1}.o $B/decode_batch.o $B/gemm_tc.o -cudart static
echo genvc_b200/libgenvc_diag${DIAG_MODE:-1}.so

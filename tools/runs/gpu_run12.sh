#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run12.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/prof_attn.py gemm
M3R_GEMM_BN=128 run python tools/prof_attn.py gemm
M3R_GEMM_BN=256 run python tools/prof_attn.py gemm
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|gemm )" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

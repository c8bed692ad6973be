#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run13.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-300} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
M3R_GEMM_PAIR=1 run python tools/test_gemm_pair.py
M3R_GEMM_PAIR=1 run python tools/prof_attn.py gemm
M3R_GEMM_PAIR=0 run python tools/prof_attn.py gemm
M3R_GEMM_PAIR=1 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_GEMM_PAIR=0 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|gemm |torch|pair|Traceback|RuntimeError)" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG

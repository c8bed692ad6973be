#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run14.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
M3R_GEMM_PAIR=0 run python tools/prof_attn.py gemm
M3R_GEMM_PAIR=2 run python tools/prof_attn.py gemm
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_GEMM_PAIR=0 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|gemm )" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG

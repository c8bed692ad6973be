#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run9.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/debug_attn.py
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python tools/prof_attn.py attn
M3R_ATTN_POLY=0 run python tools/prof_attn.py attn
run python tools/measure_parity.py
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn |224|512)" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

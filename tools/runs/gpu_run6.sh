#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run6.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/debug_attn.py
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python tools/prof_attn.py attn
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_SIDE_STREAM=0 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn|default|forced|B=1|n=1|Nm=)" $LOG | cut -c1-300
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"per_category".*"job_frac_of_peak": [0-9.]*' $LOG | cut -c1-600

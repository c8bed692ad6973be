#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run16.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python bench.py --steps 5 --warmup 3
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR)" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG
grep -o '"cpu_baseline": {[^}]*}' $LOG

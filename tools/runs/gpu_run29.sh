#!/bin/bash
# run 29: o_done wait deferred behind the exponentials: tests, attention timing, bench
mkdir -p gpurun_out; LOG=gpurun_out/run29.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=300 run python tools/prof_attn.py attn
TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|attn)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

#!/bin/bash
# run 22: closed-form tile map + coalesced split merge: tests, sweep, bench, then the traced build
mkdir -p gpurun_out; LOG=gpurun_out/run22.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=300 run python tools/prof_attn.py attn sweep
TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
echo "=== traced build" >> $LOG
M3R_ATTN_TRACE=1 timeout 600 python -m must3r_b200.build >> $LOG 2>&1
TMO=120 run python tools/trace_attn.py 7680 1 1 4
TMO=120 run python tools/trace_attn.py 768 1 1 1
TMO=120 run python tools/trace_attn.py 768 1 1 2
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|attn|sweep)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

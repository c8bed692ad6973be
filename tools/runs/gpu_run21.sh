#!/bin/bash
# run 21: memory_mode norm_y/raw tests + in-kernel timeline of the one-view attention launches
mkdir -p gpurun_out; LOG=gpurun_out/run21.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=120 run python tools/trace_attn.py 7680 1 1 4
TMO=120 run python tools/trace_attn.py 7680 1 1 2
TMO=120 run python tools/trace_attn.py 7680 1 1 1
TMO=120 run python tools/trace_attn.py 768 1 1 1
TMO=120 run python tools/trace_attn.py 768 1 1 2
TMO=120 run python tools/trace_attn.py 15360 20 2 1
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR)" $LOG | cut -c1-250 | head -40

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run5.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/debug_attn.py
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python tools/prof_attn.py attn
run python bench.py --steps 5 --warmup 3
echo "=== ncu attn full" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 -o gpurun_out/attn_v3_r01 python tools/prof_attn.py attn --once > gpurun_out/ncu_attn.log 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn|default|forced|B=1|n=1|Nm=|\{)" $LOG | cut -c1-900

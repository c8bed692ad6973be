#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run10.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python tools/prof_step.py
run python tools/prof_attn.py attn gemm
run python bench.py --steps 5 --warmup 3
echo "=== ncu step launch list" >> $LOG
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r01.csv python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
echo "--- exit $?" >> $LOG
echo "=== ncu attn full" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 -o gpurun_out/attn_v4_r01 python tools/prof_attn.py attn --once > gpurun_out/ncu_attn.log 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn |gemm |update step|render \()" $LOG | cut -c1-250
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG

#!/bin/bash
# GPU call 2: model-level parity, smoke, bench, launch list.
mkdir -p gpurun_out
LOG=gpurun_out/run2.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/measure_parity.py
run python __graft_entry__.py smoke
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python bench.py --steps 3 --warmup 3 --dtype bf16
run python bench.py --steps 3 --warmup 3 --dtype fp16 --no-cpu-baseline
run python bench.py --impl reference --steps 1 --warmup 1
echo "=== ncu launch list" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 1 --warmup 3 --views 20 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|smoke|224|512|\{)" $LOG | cut -c1-1500

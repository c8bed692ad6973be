#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run19.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python bench.py --steps 5 --warmup 3
TMO=400 run compute-sanitizer --tool memcheck --print-limit 20 python tools/san_check.py
TMO=400 run compute-sanitizer --tool racecheck --print-limit 20 python tools/san_check.py
echo "=== ncu launch list of the whole job" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_job_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "--- exit $?" >> $LOG
echo "=== ncu gemm pair + attn full" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -c 2 -o gpurun_out/gemm_pair_r01 python tools/prof_attn.py gemm --once > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 1 -o gpurun_out/attn_final_r01 python tools/prof_attn.py attn --once > gpurun_out/ncu_attn.log 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|=========|sanitizer)" $LOG | cut -c1-200 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

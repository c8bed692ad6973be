#!/bin/bash
# run 32: coalesced GEMM epilogue (smem transpose): tests, bench, GEMM timeline, big-GEMM timing
mkdir -p gpurun_out; LOG=gpurun_out/run32.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=300 run python tools/prof_attn.py gemm
echo "=== traced build" >> $LOG
M3R_TRACE=1 timeout 900 python -m must3r_b200.build >> $LOG 2>&1
TMO=120 run python tools/trace_gemm.py 768 768 768
TMO=120 run python tools/trace_gemm.py 768 768 3072
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|gemm )" $LOG | cut -c1-250 | head -60
grep -A9 "warm L2" $LOG | cut -c1-160
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG
grep -o '"clocks": {[^}]*}' $LOG

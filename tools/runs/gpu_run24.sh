#!/bin/bash
# run 24: static-weight prefetch before the PDL wait: tests, small-GEMM timing, bench
mkdir -p gpurun_out; LOG=gpurun_out/run24.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=300 run python tools/prof_attn.py small
TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_PDL=0 TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|small|layernorm)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

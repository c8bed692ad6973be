#!/bin/bash
# run 27: column-split softmax (2 warpgroups per query tile): tests, A/B timing, bench A/B
mkdir -p gpurun_out; LOG=gpurun_out/run27.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k attention
for CS in 1 2; do
  echo "=== M3R_ATTN_CS=$CS" >> $LOG
  M3R_ATTN_CS=$CS timeout 200 python tools/prof_attn.py attn 2>&1 | grep -v Warn >> $LOG
done
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
M3R_ATTN_CS=1 TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_ATTN_CS=2 TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|attn)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

#!/bin/bash
# run 20: read-back stream + e2e warm-up check, ncu of the update cross-attention (QT=1 + splits)
mkdir -p gpurun_out; LOG=gpurun_out/run20.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=600 run python bench.py --steps 5 --warmup 3
TMO=300 run python tools/prof_attn.py upd sweep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 -o gpurun_out/attn_upd_r01 python tools/prof_attn.py upd --once > gpurun_out/ncu_upd.log 2>&1
echo "--- ncu exit $?" >> $LOG
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|attn|sweep)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

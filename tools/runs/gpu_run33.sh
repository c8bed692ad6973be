#!/bin/bash
# run 33: final-state verification + streaming schedule A/B (host label shadow)
mkdir -p gpurun_out; LOG=gpurun_out/run33.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=300 run python tools/bench_stream.py 60 25
TMO=600 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|host label|device masks|identical)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG

#!/bin/bash
# run 30: round-end style verification: full GPU tests, smoke(), default bench (with CPU baseline), reference arm
mkdir -p gpurun_out; LOG=gpurun_out/run30.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=900 run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
TMO=300 run python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=600 run python bench.py
TMO=400 run python bench.py --impl reference --steps 1 --warmup 0
grep -E "^(===|--- |[0-9]+ (passed|failed)|FAILED|ERROR|smoke)" $LOG | cut -c1-250 | head -60
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus"' $LOG
grep -o '"e2e": {[^}]*}' $LOG
grep -o '"cpu_baseline": {[^}]*}' $LOG
grep -o '"clocks": {[^}]*}' $LOG

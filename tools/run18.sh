#!/bin/bash
# round 2, run 18: look-ahead variants (chain on a high-priority stream; SM budget / batch sweep)
mkdir -p gpurun_out; LOG=gpurun_out/r2_run18.log; : > $LOG
for V in "64 4" "32 4" "96 4" "64 2" "148 4"; do
  set -- $V
  echo "=== M3R_LOOKAHEAD_SMS=$1 M3R_LOOKAHEAD_BATCH=$2 (chain on a high-priority stream)" >> $LOG
  M3R_LOOKAHEAD_SMS=$1 M3R_LOOKAHEAD_BATCH=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-records --no-cpu-baseline --no-parity --encoder-mode engine 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'],1), 'views/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1))
    elif 'Error' in l or 'error' in l: print(l.strip()[:300])
" >> $LOG
done
cat $LOG

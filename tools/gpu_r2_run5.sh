#!/bin/bash
# round 2, run 5: full single-GPU bench (all records, parity, CPU baseline) + step profile after the emit reorder + full gpu pytest
mkdir -p gpurun_out; LOG=gpurun_out/r2_run5.log; : > $LOG
timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step exit $?" >> $LOG
timeout 1200 python bench.py --steps 5 --warmup 3 >> $LOG 2>&1
echo "--- bench 1 GPU (all records) exit $?" >> $LOG
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- full gpu pytest exit $?" >> $LOG
tail -40 $LOG | cut -c1-6000

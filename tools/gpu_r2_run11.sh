#!/bin/bash
# round 2, run 11: whole GPU suite (incl. keyframe scoring), smoke, default bench
mkdir -p gpurun_out; LOG=gpurun_out/r2_run11.log; : > $LOG
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- full gpu pytest exit $?" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" >> $LOG 2>&1
echo "--- smoke exit $?" >> $LOG
timeout 1200 python bench.py >> $LOG 2>&1
echo "--- default bench exit $?" >> $LOG
tail -30 $LOG | cut -c1-1500

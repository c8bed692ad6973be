#!/bin/bash
# round 2, run 15: per-(device,stream) scratch state: whole GPU suite + smoke + two host threads on two streams
mkdir -p gpurun_out; LOG=gpurun_out/r2_run15.log; : > $LOG
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- full gpu pytest exit $?" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" >> $LOG 2>&1
echo "--- smoke exit $?" >> $LOG
timeout 300 python tools/two_threads.py >> $LOG 2>&1
echo "--- two host threads exit $?" >> $LOG
tail -14 $LOG | cut -c1-400

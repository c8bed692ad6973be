#!/bin/bash
# round 2, run 14: new single-GPU tests (attention state export / merge, peer primitives), quick bench sanity
mkdir -p gpurun_out; LOG=gpurun_out/r2_run14.log; : > $LOG
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k "export_and_merge or peer_primitives or peer" >> $LOG 2>&1
echo "--- new ops tests exit $?" >> $LOG
timeout 400 python bench.py --steps 5 --warmup 3 --no-records --no-cpu-baseline >> $LOG 2>&1
echo "--- bench exit $?" >> $LOG
tail -12 $LOG | cut -c1-2500

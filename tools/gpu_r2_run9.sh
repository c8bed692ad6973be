#!/bin/bash
# round 2, run 9: split-K in the LN-emitting GEMM; step profile; bench
mkdir -p gpurun_out; LOG=gpurun_out/r2_run9.log; : > $LOG
timeout 600 python -m pytest tests/test_fused_ln_gpu.py -q --maxfail=5 --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- fused pytest exit $?" >> $LOG
timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step exit $?" >> $LOG
M3R_KSPLIT=0 timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step (M3R_KSPLIT=0) exit $?" >> $LOG
timeout 600 python -m pytest tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- model tests exit $?" >> $LOG
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-records --no-parity >> $LOG 2>&1
echo "--- bench exit $?" >> $LOG
tail -40 $LOG | cut -c1-700

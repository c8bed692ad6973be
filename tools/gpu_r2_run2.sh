#!/bin/bash
# round 2, run 2: folded-LN / merged-GEMM / emit / grouped-append decoder: unit tests, whole GPU suite, parity vs reference, bench
mkdir -p gpurun_out; LOG=gpurun_out/r2_run2.log; : > $LOG
timeout 600 python -m pytest tests/test_fused_ln_gpu.py -q --maxfail=8 --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- fused pytest exit $?" >> $LOG
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --no-header -p no:cacheprovider --deselect tests/test_fused_ln_gpu.py >> $LOG 2>&1
echo "--- full gpu pytest exit $?" >> $LOG
timeout 400 python tools/ref_gpu_parity.py c2 c3 >> $LOG 2>&1
echo "--- ref parity exit $?" >> $LOG
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
echo "--- bench exit $?" >> $LOG
M3R_EMIT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
echo "--- bench (M3R_EMIT=0) exit $?" >> $LOG
tail -80 $LOG | cut -c1-300

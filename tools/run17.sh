#!/bin/bash
# round 2, run 17: encoder look-ahead next to the decoder chain: correctness (engine tests) + C3 A/B
mkdir -p gpurun_out; LOG=gpurun_out/r2_run17.log; : > $LOG
timeout 600 python -m pytest tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider -k "engine or chains or stream" >> $LOG 2>&1
echo "--- engine tests exit $?" >> $LOG
timeout 400 python bench.py --steps 5 --warmup 3 --no-records --no-cpu-baseline >> $LOG 2>&1
echo "--- bench (engine-managed encoding, look-ahead) exit $?" >> $LOG
timeout 400 python bench.py --steps 5 --warmup 3 --no-records --no-cpu-baseline --no-parity --encoder-mode precomputed >> $LOG 2>&1
echo "--- bench (precomputed features, as before) exit $?" >> $LOG
M3R_LOOKAHEAD=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-records --no-cpu-baseline --no-parity >> $LOG 2>&1
echo "--- bench (engine-managed, M3R_LOOKAHEAD=0: per-step encoding like the reference) exit $?" >> $LOG
tail -12 $LOG | cut -c1-700

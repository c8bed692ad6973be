#!/bin/bash
# round 2, run 3: where does the one-view step spend its time after the restructure?
mkdir -p gpurun_out; LOG=gpurun_out/r2_run3.log; : > $LOG
timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step exit $?" >> $LOG
M3R_EMIT=0 timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step (M3R_EMIT=0) exit $?" >> $LOG
M3R_SIDE_STREAM=0 timeout 200 python tools/prof_step.py >> $LOG 2>&1
echo "--- prof_step (M3R_SIDE_STREAM=0) exit $?" >> $LOG
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-records --no-parity >> $LOG 2>&1
echo "--- bench exit $?" >> $LOG
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_update_step.csv python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
echo "--- ncu step exit $?" >> $LOG
timeout 600 python -m pytest tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider -k "chains or inplace or stream_schedule or errors" -s >> $LOG 2>&1
echo "--- new model tests exit $?" >> $LOG
timeout 100 python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k "errors_are_loud" >> $LOG 2>&1
echo "--- ops errors test exit $?" >> $LOG
tail -70 $LOG | cut -c1-400

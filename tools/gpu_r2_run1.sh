#!/bin/bash
# round 2, run 1: baselines on the round-1 kernels + seam tests + reference-on-GPU parity at the benched configs
mkdir -p gpurun_out; LOG=gpurun_out/r2_run1.log; : > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $LOG 2>&1
timeout 300 python -m pytest tests/test_reference_seam_gpu.py -q -x --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- seam pytest exit $?" >> $LOG
timeout 400 python tools/ref_gpu_parity.py c2 c3 >> $LOG 2>&1
echo "--- ref parity exit $?" >> $LOG
timeout 300 python tools/prof_attn.py longmem >> $LOG 2>&1
echo "--- longmem exit $?" >> $LOG
timeout 200 python bench.py --views 100 --steps 2 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
echo "--- bench 100 views exit $?" >> $LOG
timeout 200 python tools/bench_stream.py 300 25 >> $LOG 2>&1
echo "--- stream exit $?" >> $LOG
tail -60 $LOG | cut -c1-400

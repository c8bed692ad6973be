#!/bin/bash
# run 23: split heuristic check + in-kernel timeline of the one-view GEMMs
mkdir -p gpurun_out; LOG=gpurun_out/run23.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TMO=600 run python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider
TMO=300 run python tools/prof_attn.py small --warm
echo "=== traced build" >> $LOG
M3R_TRACE=1 timeout 600 python -m must3r_b200.build >> $LOG 2>&1
TMO=120 run python tools/trace_gemm.py 768 768 768
TMO=120 run python tools/trace_gemm.py 768 3072 768
TMO=120 run python tools/trace_gemm.py 768 768 3072
TMO=120 run python tools/trace_gemm.py 768 768 768 128
TMO=120 run python tools/trace_attn.py 7680 1
cat $LOG | cut -c1-200 | grep -v "^$" | head -150

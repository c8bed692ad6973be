#!/bin/bash
# run 28: MMA-thread stamps: when is Q K^T (j+1) issued relative to the softmax needing it?
mkdir -p gpurun_out; LOG=gpurun_out/run28.log; : > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
M3R_TRACE=1 timeout 900 python -m must3r_b200.build >> $LOG 2>&1
TMO=120 run python tools/trace_attn.py 768 1 1 1
TMO=120 run python tools/trace_attn.py 7680 1 1 4
TMO=120 run python tools/trace_attn.py 15360 20 2 1
cat $LOG | cut -c1-260

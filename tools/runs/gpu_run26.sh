#!/bin/bash
# run 26: polynomial-exp fraction sweep on the one-view (QT=1, two CTAs per SM) attention shapes
mkdir -p gpurun_out; LOG=gpurun_out/run26.log; : > $LOG
for P in 0 1 2 3; do
  echo "=== M3R_ATTN_POLY=$P" >> $LOG
  M3R_ATTN_POLY=$P timeout 200 python tools/prof_attn.py attn 2>&1 | grep -v Warn >> $LOG
done
cat $LOG

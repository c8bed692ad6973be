#!/bin/bash
mkdir -p gpurun_out; LOG=gpurun_out/r2_run8.log; : > $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_context_parallel.py >> $LOG 2>&1
echo "--- check_context_parallel exit $?" >> $LOG
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" $LOG | tail -30 | cut -c1-600

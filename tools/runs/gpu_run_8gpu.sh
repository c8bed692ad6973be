#!/bin/bash
# 8-GPU sanity of bench.py (the driver's scaling run ends at N=8)
N=${1:-8}
mkdir -p gpurun_out
LOG=gpurun_out/run_${N}gpu.log
: > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench N=$N" >> $LOG
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|Traceback|RuntimeError|.*Error)" $LOG | cut -c1-300
grep -o "\"value\": [0-9.]*, \"unit\": \"views/s\", \"n_gpus\": $N" $LOG
grep -o '"e2e": {[^}]*}' $LOG
grep -o '"ms_per_step": [0-9.]*' $LOG | head -2

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run_2gpu_c.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-240} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run $TR --master-port 29511 tools/check_sharded.py
M3R_FUSED_GATHER=0 run $TR --master-port 29512 tools/check_sharded.py
run $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3
M3R_FUSED_GATHER=0 run $TR --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3
grep -E "^(===|--- exit|rank|Traceback|RuntimeError|.*Error)" $LOG | cut -c1-300
grep -o '"value": [0-9.]*, "unit": "views/s", "n_gpus": 2' $LOG
grep -o '"e2e": {[^}]*}' $LOG

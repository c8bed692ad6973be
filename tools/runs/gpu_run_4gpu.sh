#!/bin/bash
# N-GPU sanity of the sharded schedule + bench (N from $1, default 4)
N=${1:-4}
mkdir -p gpurun_out
LOG=gpurun_out/run_${N}gpu.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-240} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
TMO=200 run $TR --master-port 29511 tools/check_sharded.py
TMO=300 run $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|rank|Traceback|RuntimeError|.*Error)" $LOG | cut -c1-300
grep -o "\"value\": [0-9.]*, \"unit\": \"views/s\", \"n_gpus\": $N" $LOG
grep -o '"e2e": {[^}]*}' $LOG

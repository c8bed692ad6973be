#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run_2gpu_b.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-240} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3
grep -E "^(===|--- exit|rank|\{)" $LOG | cut -c1-2500

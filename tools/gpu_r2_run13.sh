#!/bin/bash
# round 2, run 13 (8 GPUs): final multi-GPU validation: checks, C4 A/B (fused with/without the append CTA cap, NCCL gather), C5 on 8 GPUs
mkdir -p gpurun_out; LOG=gpurun_out/r2_run13.log; : > $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 tools/check_sharded.py >> $LOG 2>&1
echo "--- check_sharded 8 GPUs exit $?" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/check_context_parallel.py >> $LOG 2>&1
echo "--- check_context_parallel 8 GPUs exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 5 --warmup 3 >> $LOG 2>&1
echo "--- bench 8 GPUs (default: fused, cap) exit $?" >> $LOG
M3R_APPEND_CAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --steps 5 --warmup 3 --no-parity --no-records >> $LOG 2>&1
echo "--- bench 8 GPUs (fused, M3R_APPEND_CAP=0) exit $?" >> $LOG
M3R_FUSED_GATHER=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 5 --warmup 3 --no-parity --no-records >> $LOG 2>&1
echo "--- bench 8 GPUs (NCCL all-gather path) exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --config c5 --steps 1 --warmup 3 --no-records >> $LOG 2>&1
echo "--- bench c5 8 GPUs exit $?" >> $LOG
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" $LOG | tail -20 | cut -c1-1200

#!/bin/bash
# round 2, run 6 (8 GPUs): sharded schedule check + C4 bench at 8 and 4 GPUs
mkdir -p gpurun_out; LOG=gpurun_out/r2_run6.log; : > $LOG
nvidia-smi -L >> $LOG 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 tools/check_sharded.py >> $LOG 2>&1
echo "--- check_sharded 8 GPUs exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 5 --warmup 3 >> $LOG 2>&1
echo "--- bench 8 GPUs exit $?" >> $LOG
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --no-parity >> $LOG 2>&1
echo "--- bench 4 GPUs exit $?" >> $LOG
M3R_FUSED_GATHER=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 3 --warmup 3 --no-parity --no-records >> $LOG 2>&1
echo "--- bench 8 GPUs NCCL all-gather path exit $?" >> $LOG
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" $LOG | tail -30 | cut -c1-1500

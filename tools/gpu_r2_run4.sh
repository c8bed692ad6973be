#!/bin/bash
# round 2, run 4 (2 GPUs): sharded schedule with the device-side barrier, ragged split, C4 at N=2; single-GPU bench with records
mkdir -p gpurun_out; LOG=gpurun_out/r2_run4.log; : > $LOG
nvidia-smi -L >> $LOG 2>&1
for F in 1 0; do
M3R_FUSED_GATHER=$F timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$F tools/check_sharded.py >> $LOG 2>&1
echo "--- check_sharded fused=$F exit $?" >> $LOG
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 >> $LOG 2>&1
echo "--- bench 2 GPUs exit $?" >> $LOG
timeout 900 python bench.py --steps 5 --warmup 3 >> $LOG 2>&1
echo "--- bench 1 GPU (all records) exit $?" >> $LOG
tail -40 $LOG | cut -c1-3000

#!/bin/bash
# round 2, run 7 (2 GPUs): context-parallel cross-attention check + C5 stream on 2 GPUs + C4 on 2 GPUs after the append-GEMM cap
mkdir -p gpurun_out; LOG=gpurun_out/r2_run7.log; : > $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_context_parallel.py >> $LOG 2>&1
echo "--- check_context_parallel exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --config c5 --steps 1 --warmup 3 --no-records >> $LOG 2>&1
echo "--- bench c5 2 GPUs exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity >> $LOG 2>&1
echo "--- bench c4 2 GPUs exit $?" >> $LOG
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k "attention" >> $LOG 2>&1
echo "--- attention tests exit $?" >> $LOG
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" $LOG | tail -30 | cut -c1-1800

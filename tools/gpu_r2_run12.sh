#!/bin/bash
# round 2, run 12 (2 GPUs): staged coalesced peer stores: correctness (check_sharded, pytest 2-GPU tests) + C4 at N=2
mkdir -p gpurun_out; LOG=gpurun_out/r2_run12.log; : > $LOG
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_keyframes_gpu.py tests/test_reference_seam_gpu.py -q -x --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- 2-GPU pytest exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity >> $LOG 2>&1
echo "--- bench c4 2 GPUs exit $?" >> $LOG
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" $LOG | tail -12 | cut -c1-1500

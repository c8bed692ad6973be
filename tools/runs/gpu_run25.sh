#!/bin/bash
# run 25: ncu --set full of the dominant small kernels (gemm_kernel<64> one-view GEMM, attn_kernel<QT=1> update CA)
mkdir -p gpurun_out; LOG=gpurun_out/run25.log; : > $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o gpurun_out/gemm64_r01 python tools/prof_attn.py g64 --once > gpurun_out/ncu_g64.log 2>&1
echo "--- ncu g64 exit $?" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 -o gpurun_out/attn_upd2_r01 python tools/prof_attn.py upd --once > gpurun_out/ncu_upd2.log 2>&1
echo "--- ncu upd exit $?" >> $LOG
cat $LOG

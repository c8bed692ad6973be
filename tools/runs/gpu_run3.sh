#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run3.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/diag_kernels.py attn
run python -m pytest tests/test_ops_gpu.py -q -x -m gpu --no-header -p no:cacheprovider -k "attention"
run python tools/prof_attn.py attn gemm
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
echo "=== ncu attn full" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 4 -o gpurun_out/attn_r01 python tools/prof_attn.py attn --once > gpurun_out/ncu_attn.log 2>&1
echo "--- exit $?" >> $LOG
echo "=== ncu gemm full" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 6 -o gpurun_out/gemm_r01 python tools/prof_attn.py gemm --once > gpurun_out/ncu_gemm.log 2>&1
echo "--- exit $?" >> $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn|gemm|\{)" $LOG | cut -c1-1200

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run4.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/diag_kernels.py attn
run python -m pytest tests/test_ops_gpu.py -q -x -m gpu --no-header -p no:cacheprovider
run python tools/prof_attn.py attn sweep
run python -m pytest tests -q -x -m gpu --no-header -p no:cacheprovider --deselect tests/test_ops_gpu.py
run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
M3R_PDL=0 run python bench.py --steps 5 --warmup 3 --no-cpu-baseline
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR|attn|gemm|sweep|\{)" $LOG | cut -c1-700

#!/bin/bash
# Run on the GPU box (under gpurun): isolated, time-bounded bring-up of every kernel family.
# Each step runs in its own process with a timeout so one trap / hang does not hide the others.
mkdir -p gpurun_out
LOG=gpurun_out/gpu_check.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
run() { echo "=== $*" >> $LOG; timeout 300 "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/diag_kernels.py gemm
run python tools/diag_kernels.py attn
for t in test_gemm_plain test_gemm_large_persistent test_gemm_epilogues test_gemm_rope_epilogue test_rope_2d_curope_contract \
         test_layernorm test_self_attention test_memory_cross_attention test_render_cross_attention_long_memory \
         test_patch_embed_path test_unpatchify_and_postprocess test_errors_are_loud; do
  run python -m pytest tests/test_ops_gpu.py -q -x -m gpu -k "$t" --no-header -p no:cacheprovider
done
tail -5 $LOG
grep -E "^(===|--- exit|[0-9]+ (passed|failed)|FAILED|ERROR)" $LOG | tail -80

#!/bin/bash
# run 34 (last GPU seconds of the round): engine-on-CUDA tests with the tensor-typed label shadow + compute_cam on CUDA
mkdir -p gpurun_out; LOG=gpurun_out/run34.log; : > $LOG
timeout 70 python -m pytest tests/test_model_gpu.py -q -x -k "engine" --no-header -p no:cacheprovider >> $LOG 2>&1
echo "--- pytest exit $?" >> $LOG
timeout 40 python - >> $LOG 2>&1 <<'PY'
import torch
from must3r_b200 import engine
pm = torch.randn(1, 3, 32, 48, 7, device="cuda")
out = engine.postprocess(pm, "norm_exp", compute_cam=True)
ref = engine.postprocess(pm.cpu(), "norm_exp", compute_cam=True)
print("compute_cam cuda vs cpu: focal", float((out["focal"].cpu() - ref["focal"]).abs().max()),
      "c2w", float((out["c2w"].cpu() - ref["c2w"]).abs().max()))
PY
echo "--- cam exit $?" >> $LOG
tail -12 $LOG | cut -c1-200

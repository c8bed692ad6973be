#!/bin/bash
# round 2, run 10: ncu evidence (full set on the dominant kernels, launch list of the whole C3 job)
mkdir -p gpurun_out; LOG=gpurun_out/r2_run10.log; : > $LOG
for T in emit emitk merged grouped ca1 carender; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_kernel" -s 1 -c 2 -f -o gpurun_out/r02_$T python tools/ncu_targets.py $T >> $LOG 2>&1
  echo "--- ncu $T exit $?" >> $LOG
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 2400 --csv --log-file gpurun_out/r02_launches_job.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-records --no-parity > gpurun_out/ncu_job.log 2>&1
echo "--- ncu job launch list exit $?" >> $LOG
ls -la gpurun_out/r02_* >> $LOG
tail -15 $LOG

#!/bin/bash
# Round-2 evidence pass on one B200 (final kernels): per-CTA phase traces, ncu launch lists at both batch sizes, ncu --set full
# of one step's conv_tc2 launches and of a conv_pm2 launch, compute-sanitizer memcheck / racecheck / synccheck, the
# UNet-only batch sweep (BASELINE configs[3]).      bash tools/gpu_r2_evidence.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python tools/tc_trace.py 8190 f16x3 > gpurun_out/${TAG}_cta_phases_8190.txt 2>&1
EDMP_NO_RUNS=1 timeout 300 python tools/tc_trace.py 8190 f16x3 > gpurun_out/${TAG}_cta_phases_8190_per_layer.txt 2>&1
timeout 300 python tools/tc_trace.py 1020 f16x3 > gpurun_out/${TAG}_cta_phases_1020.txt 2>&1
for ROWS in 8190 1020; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${ROWS}.csv \
      python tools/ncu_target.py $ROWS 4 f16x3 > gpurun_out/${TAG}_ncu_launches_${ROWS}.log 2>&1
done
# one whole step's conv_tc2 launches (7 at 8190 rows: 2 single-CTA runs, 3 CTA-pair runs, 2 single resampling layers), second step
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 7 -c 7 -f -o gpurun_out/${TAG}_tc2_full \
    python tools/ncu_target.py 8190 3 f16x3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pm2 -s 16 -c 3 -f -o gpurun_out/${TAG}_pm2_full \
    python tools/ncu_target.py 8190 2 f16x3 > gpurun_out/${TAG}_ncu_full_pm2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tail -s 2 -c 2 -f -o gpurun_out/${TAG}_tail_full \
    python tools/ncu_target.py 8190 4 f16x3 > gpurun_out/${TAG}_ncu_full_tail.log 2>&1
for TOOL in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $TOOL python tools/sanitize_target.py 130 > gpurun_out/${TAG}_${TOOL}_130rows.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_${TOOL}_130rows.log
  EDMP_CG2=1 timeout 900 compute-sanitizer --tool $TOOL python tools/sanitize_target.py 300 > gpurun_out/${TAG}_${TOOL}_300rows_cta_pairs.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_${TOOL}_300rows_cta_pairs.log
  tail -3 gpurun_out/${TAG}_${TOOL}_130rows.log gpurun_out/${TAG}_${TOOL}_300rows_cta_pairs.log
done
timeout 900 python tools/bench_unet_sweep.py f16x3 gpurun_out/${TAG}_unet_batch_sweep.txt > gpurun_out/${TAG}_sweep.log 2>&1; cat gpurun_out/${TAG}_unet_batch_sweep.txt
ls -la gpurun_out/${TAG}_*

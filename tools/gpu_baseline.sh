#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_baseline.sh <tag> [precision] [rows]
TAG=${1:-run}; PREC=${2:-f16x3}; ROWS=${3:-8190}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --precision $PREC --ops-out gpurun_out/${TAG}_ops.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/ncu_target.py $ROWS 4 $PREC > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 30 -c 4 -f -o gpurun_out/${TAG}_tc2_full \
    python tools/ncu_target.py $ROWS 2 $PREC > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench.json

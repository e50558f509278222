#!/bin/bash
# per-CTA phase traces with the trace-mode ablation knob (EDMP_ABLATE: 1 = epilogues idle, 2 = no MMAs issued)
TAG=${1:-abl}; ROWS=${2:-"8190 1020"}
mkdir -p gpurun_out
for R in $ROWS; do
  for A in 0 1 2; do
    EDMP_ABLATE=$A timeout 300 python tools/tc_trace.py $R f16x3 > gpurun_out/${TAG}_${R}_a${A}.txt 2>&1
  done
done
for f in gpurun_out/${TAG}_*_a1.txt; do tail -n 3 $f; done

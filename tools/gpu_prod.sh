#!/bin/bash
# operand-producer thread count A/B (EDMP_PRODUCERS), with and without the lean issuer
mkdir -p gpurun_out
EDMP_PRODUCERS=3 EDMP_MMA_LEAN=1 EDMP_TEST_PRECISIONS=f16x3 timeout 600 python -m pytest tests -m gpu -x -q -k "unet" 2>&1 | tail -3
EDMP_PRODUCERS=2 EDMP_TEST_PRECISIONS=f16x3 timeout 600 python -m pytest tests -m gpu -x -q -k "unet_matches or headline or cta_pairs" 2>&1 | tail -3
bash tools/gpu_ab_env2.sh p1 ""
bash tools/gpu_ab_env2.sh p2 "EDMP_PRODUCERS=2"
bash tools/gpu_ab_env2.sh p3 "EDMP_PRODUCERS=3 EDMP_MMA_LEAN=1"
bash tools/gpu_ab_env2.sh p2l "EDMP_PRODUCERS=2 EDMP_MMA_LEAN=1"

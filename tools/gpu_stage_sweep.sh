#!/bin/bash
# operand-feed experiment: per-launch times of an 8190-row forward with capped operand stage rings
TAG=${1:-stg}
mkdir -p gpurun_out
for CFG in "9 9" "2 9" "9 1" "2 1" "4 1"; do
  set -- $CFG
  EDMP_MAX_A_STAGES=$1 EDMP_MAX_B_STAGES=$2 timeout 300 python bench.py --precision f16x3 --rows-per-guide 819 --quick --steps 2 --ops-out gpurun_out/${TAG}_ops_a$1_b$2.txt > gpurun_out/${TAG}_bench_a$1_b$2.json 2> gpurun_out/${TAG}_a$1_b$2.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_a$1_b$2.json'))
print('A<=$1 B<=$2', 'value', round(d['value'],1), 'unet', d['unet']['ms_per_forward'], d['roofline']['by_kernel_ms'])
"
  grep "layers)\|down_samplers.4.down.3\|up_samplers.0.up.3" gpurun_out/${TAG}_ops_a$1_b$2.txt
done

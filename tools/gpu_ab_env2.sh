#!/bin/bash
# A/B of environment knobs: per-launch times + in-stream value at both batch sizes.  usage: gpu_ab_env2.sh <tag> "<ENV=1 ...>" [rows-per-guide list]
TAG=${1:-ab}; ENVS=${2:-""}; RPGS=${3:-"819 102"}
mkdir -p gpurun_out
for RPG in $RPGS; do
  env $ENVS timeout 300 python bench.py --precision f16x3 --rows-per-guide $RPG --quick --steps 2 --ops-out gpurun_out/${TAG}_ops_${RPG}.txt > gpurun_out/${TAG}_bench_${RPG}.json 2> gpurun_out/${TAG}_${RPG}.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_${RPG}.json'))
print('$ENVS', 'rows', d['config']['rows_per_gpu'], 'value', round(d['value'],1), 'unet', round(d['unet']['ms_per_forward'],3), d['roofline']['by_kernel_ms'], d['clocks']['sm_mhz'])
" || tail -3 gpurun_out/${TAG}_${RPG}.err
done

#!/bin/bash
# A/B of one environment knob on the default bench workload:  bash tools/gpu_ab_env.sh <tag> <VAR> <value> [<value> ...]
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for V in "$@"; do
  if [ "$V" = "-" ]; then unset $VAR; else export $VAR=$V; fi
  timeout 600 python bench.py --quick --steps 2 --ops-out gpurun_out/${TAG}_${V}_ops.txt > gpurun_out/${TAG}_${V}.json 2> gpurun_out/${TAG}_${V}.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_${V}.json'))
print('$VAR=$V', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'unet ms', round(d['unet']['ms_per_forward'],4), d['roofline']['by_kernel_ms'])
" || tail -5 gpurun_out/${TAG}_${V}.err
done

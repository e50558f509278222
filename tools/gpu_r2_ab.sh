#!/bin/bash
# A/B of environment knobs: bash tools/gpu_r2_ab.sh <tag> <rows-per-guide> "<ENV=V ...>" "<ENV=V ...>" ...   ("-" = no knob)
TAG=$1; RPG=$2; shift 2
mkdir -p gpurun_out
i=0
for KNOBS in "$@"; do
  i=$((i+1))
  if [ "$KNOBS" = "-" ]; then E=""; else E="$KNOBS"; fi
  env $E timeout 600 python bench.py --rows-per-guide $RPG --quick --steps 3 --ops-out gpurun_out/${TAG}_${i}_ops.txt > gpurun_out/${TAG}_${i}.json 2> gpurun_out/${TAG}_${i}.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_${i}.json'))
print('[$KNOBS]', 'rows', d['config']['rows_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'unet ms', round(d['unet']['ms_per_forward'],4), 'launches', d['unet']['launches'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4), 'clk', d['clocks']['sm_mhz'])
" || tail -5 gpurun_out/${TAG}_${i}.err
done

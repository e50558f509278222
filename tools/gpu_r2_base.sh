#!/bin/bash
# round-2 baseline: per-CTA traces at both batch sizes, short bench at both, per-op tables
TAG=${1:-r2base}; PREC=f16x3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python tools/tc_trace.py 1020 $PREC > gpurun_out/${TAG}_trace_1020.txt 2>&1
timeout 300 python tools/tc_trace.py 8190 $PREC > gpurun_out/${TAG}_trace_8190.txt 2>&1
for RPG in 102 819; do
  timeout 600 python bench.py --precision $PREC --rows-per-guide $RPG --quick --steps 2 --ops-out gpurun_out/${TAG}_ops_${RPG}.txt > gpurun_out/${TAG}_bench_${RPG}.json 2> gpurun_out/${TAG}_bench_${RPG}.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_${RPG}.json'))
print('rows/gpu', d['config']['rows_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'unet', d['unet'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4), d['clocks'])
" || tail -5 gpurun_out/${TAG}_bench_${RPG}.err
done

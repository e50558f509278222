#!/bin/bash
# one development iteration on the GPU box: parity tests (f16x3), per-CTA trace, short bench at two batch sizes
TAG=${1:-it}; PREC=${2:-f16x3}
mkdir -p gpurun_out
EDMP_TEST_PRECISIONS=$PREC timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
timeout 300 python tools/tc_trace.py 8190 $PREC > gpurun_out/${TAG}_trace_8190.txt 2>&1
for RPG in 102 819; do
  timeout 600 python bench.py --precision $PREC --rows-per-guide $RPG --quick --steps 2 --ops-out gpurun_out/${TAG}_ops_${RPG}.txt > gpurun_out/${TAG}_bench_${RPG}.json 2> gpurun_out/${TAG}_bench_${RPG}.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_${RPG}.json'))
print('rows/gpu', d['config']['rows_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'unet', d['unet'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4))
" || tail -5 gpurun_out/${TAG}_bench_${RPG}.err
done

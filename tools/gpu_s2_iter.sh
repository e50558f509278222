#!/bin/bash
# session-2 development iteration: UNet parity tests, ablation traces at 8190 rows, short bench at both batch sizes.
# usage: bash tools/gpu_s2_iter.sh <tag> [pytest -k expression] [rows-per-guide list] [ablate modes]
TAG=${1:-it}; KEXPR=${2:-unet}; RPGS=${3:-"819 102"}; ABL=${4:-"0 1 2"}; PREC=f16x3
mkdir -p gpurun_out
EDMP_TEST_PRECISIONS=$PREC timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
for A in $ABL; do
  for R in ${ABL_ROWS:-8190}; do EDMP_ABLATE=$A timeout 300 python tools/tc_trace.py $R $PREC > gpurun_out/${TAG}_abl_${R}_a${A}.txt 2>&1; done
done
for RPG in $RPGS; do
  timeout 600 python bench.py --precision $PREC --rows-per-guide $RPG --quick --steps 2 --ops-out gpurun_out/${TAG}_ops_${RPG}.txt > gpurun_out/${TAG}_bench_${RPG}.json 2> gpurun_out/${TAG}_bench_${RPG}.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_${RPG}.json'))
print('rows/gpu', d['config']['rows_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'unet', d['unet'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4), d['roofline']['by_kernel_ms'], d['clocks'])
" || tail -5 gpurun_out/${TAG}_bench_${RPG}.err
done

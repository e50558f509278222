#!/bin/bash
# full check: the whole GPU test suite, smoke(), the default bench line (all blocks), the c3 line
TAG=${1:-full}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log
( time timeout 1200 python bench.py --ops-out gpurun_out/${TAG}_ops.txt ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').readline())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'strong',d['strong'],'\napi',d['e2e_api'],'\ngpu_baseline',d['gpu_baseline'],'\ncpu',d['cpu_baseline'],'\nroof',d['roofline']['kernel'],round(d['roofline']['frac'],4),d['clocks'])
PY
( time timeout 600 python bench.py --config c3 --no-cpu-baseline --no-api-e2e --ops-out gpurun_out/${TAG}_c3_ops.txt ) > gpurun_out/${TAG}_bench_c3.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_c3.json').readline())
print('c3 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'unet',d['unet'],'gpu_baseline',d['gpu_baseline'])
PY

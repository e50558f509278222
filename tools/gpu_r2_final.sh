#!/bin/bash
# Round-2 final pass on one B200: smoke, the whole GPU suite, the default bench line (all blocks), the configs[2] line, the
# 1020-rows/GPU shard line, both baseline arms, sanitizer on the multi-layer runs (8190-row forward).   bash tools/gpu_r2_final.sh <tag>
TAG=${1:-r2final}
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 --ops-out gpurun_out/${TAG}_ops.txt ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
( time timeout 600 python bench.py --config c3 --no-api-e2e --ops-out gpurun_out/${TAG}_c3_ops.txt ) > gpurun_out/${TAG}_bench_c3.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --rows-per-guide 102 --no-cpu-baseline --no-gpu-baseline --no-api-e2e --no-strong --ops-out gpurun_out/${TAG}_ops_1020.txt > gpurun_out/${TAG}_bench_1020.json 2>> gpurun_out/${TAG}_bench.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl torch-cuda > gpurun_out/${TAG}_bench_torch_cuda.json 2>> gpurun_out/${TAG}_bench.err
for TOOL in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $TOOL python tools/sanitize_target.py 8190 unet > gpurun_out/${TAG}_${TOOL}_8190rows_runs.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_${TOOL}_8190rows_runs.log
  tail -4 gpurun_out/${TAG}_${TOOL}_8190rows_runs.log
done
python - <<PY
import json
for f in ("bench","bench_c3","bench_1020","bench_reference","bench_torch_cuda"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_%s.json"%f) if l.startswith("{")][-1])
        print(f, "value", round(d["value"],2), "e2e", d.get("e2e",{}).get("value"), "strong", (d.get("strong") or {}).get("value"), "api", d.get("e2e_api"), "gpu_baseline", d.get("gpu_baseline"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roof", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY

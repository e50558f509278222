#!/bin/bash
# Round-end evidence pass on one B200: smoke, full GPU tests, the default bench line, the 1020-row shard line, ncu launch
# list and full captures of both conv_tc2 variants and of conv_pm2.   bash tools/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --ops-out gpurun_out/${TAG}_ops.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --rows-per-guide 102 --no-cpu-baseline --ops-out gpurun_out/${TAG}_ops_1020.txt > gpurun_out/${TAG}_bench_1020.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/ncu_target.py 8190 4 f16x3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 30 -c 4 -f -o gpurun_out/${TAG}_tc2_full \
    python tools/ncu_target.py 8190 2 f16x3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pm2 -s 8 -c 2 -f -o gpurun_out/${TAG}_pm2_full \
    python tools/ncu_target.py 8190 1 f16x3 > gpurun_out/${TAG}_ncu_full_pm2.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log; tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_1020.json | cut -c1-400; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-600

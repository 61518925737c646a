#!/bin/bash
# What the driver runs at round end, on one B200, plus the launch lists for profiles/.
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_launches_smoke.csv python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_final_reference.json 2> gpurun_out/final_ref.err
tail -c 400 gpurun_out/r02_final_reference.json
timeout 1200 python bench.py --steps 50 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/r02_final_bench.json; tail -2 gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:hnsw_search|pad_rows|merge_topk|exchange_merge' -c 60 --csv --log-file gpurun_out/r02_launches_c3s_steps.csv python bench.py --workload c3s --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r02_launches_c3s_steps.csv
timeout 600 python tools/io_probe.py 10000000 768 ip /tmp 2>&1 | tail -1

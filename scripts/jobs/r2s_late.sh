#!/bin/bash
# late round 2: refreshed other-config numbers and an ncu capture of the cluster kernel with the split barriers
python scripts/bench_configs.py > gpurun_out/r2s_configs.jsonl 2> gpurun_out/r2s_configs.err; echo "configs rc=$?"
python scripts/bench_operators.py >> gpurun_out/r2s_configs.jsonl 2>> gpurun_out/r2s_configs.err
cat gpurun_out/r2s_configs.jsonl | cut -c1-200
QGB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:qg_step_cl_kernel -s 3 -c 1 \
  -o gpurun_out/r2s_cluster256 -f python scripts/run_large_once.py 256 64 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log

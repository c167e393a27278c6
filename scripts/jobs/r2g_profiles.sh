#!/bin/bash
# ncu --set full of every kernel of the headline step (tc_fast) + the 256^2 cluster kernel; launch list of the final code
B="python bench.py --steps 3 --warmup 3 --precision tc_fast --no-library-baseline --cpu-steps 1 --e2e-steps 1"
QGB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/ncu_l.log 2>&1; echo "launch list rc=$?"
QGB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 21 -c 7 -f -o gpurun_out/r2g_conv_tc $B > gpurun_out/ncu_conv.log 2>&1; echo "conv rc=$?"
ncu --set full --clock-control none --import-source on -k qg_step_cl_kernel -s 3 -c 1 -f -o gpurun_out/r2g_cl256 python scripts/run_large_once.py 256 64 > gpurun_out/ncu_cl.log 2>&1; echo "cl rc=$?"
python bench.py > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; echo "bench rc=$?"
python scripts/bench_large.py > gpurun_out/r2g_bench_large.log 2>&1; tail -4 gpurun_out/r2g_bench_large.log

#!/bin/bash
# round-2 evidence job: GPU tests, 1-GPU bench line, launch list, ncu --set full of the kernels furthest from their roofline
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo "bench rc=$?"
B="python bench.py --steps 3 --warmup 3 --precision tc_fast --no-library-baseline --cpu-steps 1 --e2e-steps 1"
QGB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/ncu_l.log 2>&1
for k in qg_step64_kernel conv_l1_direct_kernel; do
  QGB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k $k -s 3 -c 1 -f -o gpurun_out/r2b_$k $B > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
ncu --set full --clock-control none --import-source on -k qg_step_cl_kernel -s 3 -c 1 -f -o gpurun_out/r2b_cl256 python scripts/run_large_once.py 256 64 > gpurun_out/ncu_cl.log 2>&1; echo "cl rc=$?"
python scripts/bench_large.py > gpurun_out/r2b_bench_large.log 2>&1; echo "large rc=$?"
tail -3 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_bench_large.log | tail -8

#!/bin/bash
python scripts/bench_configs.py > gpurun_out/r2j_configs.jsonl 2> gpurun_out/r2j_configs.err; echo "configs rc=$?"; cat gpurun_out/r2j_configs.jsonl
python scripts/bench_operators.py > gpurun_out/r2j_operators.log 2>&1; tail -6 gpurun_out/r2j_operators.log
python scripts/spectral_time.py 64 1024 200 2>&1 | tail -2
python scripts/gap_probe.py > gpurun_out/r2j_gap.log 2>&1; tail -8 gpurun_out/r2j_gap.log
python scripts/bench_plugin.py > gpurun_out/r2j_plugin.log 2>&1; tail -4 gpurun_out/r2j_plugin.log

#!/bin/bash
# A/B of the cluster register-FFT geometries: 256-thread CTAs (two per SM) vs 512-thread CTAs (QGB_SCL_WIDE=1)
python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c_tests.log
python scripts/bench_large.py > gpurun_out/r2c_bench_large_narrow.log 2>&1; cat gpurun_out/r2c_bench_large_narrow.log | tail -5
QGB_SCL_WIDE=1 python scripts/bench_large.py > gpurun_out/r2c_bench_large_wide.log 2>&1; cat gpurun_out/r2c_bench_large_wide.log | tail -5

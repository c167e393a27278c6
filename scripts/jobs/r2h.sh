#!/bin/bash
python -m pytest tests/test_gpu_spectral.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python scripts/bench_large.py 2>&1 | tail -4
python scripts/bench_training.py > gpurun_out/r2h_training.jsonl 2> gpurun_out/r2h_training.err; cat gpurun_out/r2h_training.jsonl; tail -3 gpurun_out/r2h_training.err

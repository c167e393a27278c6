#!/bin/bash
python -m pytest tests/test_gpu_closure.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -4
python scripts/bench_training.py 2>/dev/null | head -1
QGB_WGRAD5_WIDE=1 python scripts/bench_training.py 2>/dev/null | head -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_train_launches.csv python scripts/train_once.py 64 64 2 > gpurun_out/r2k.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2k_train_launches.csv | head -8

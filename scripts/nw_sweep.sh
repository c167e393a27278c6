#!/bin/bash
# Experiment: depth of the weight-stage ring (stages hold one tap row) vs throughput.  Rebuilds the library per setting.
cd "$(dirname "$0")/.."
for nw in 2 3 5; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DQGB_TC_NW_OVERRIDE=$nw -o pyqg_generative_b200/libqgb200.so pyqg_generative_b200/csrc/api.cu 2>&1 | grep -E "error"
  echo "NW=$nw"
  timeout 200 python bench.py --steps 20 --warmup 3 --cpu-steps 1 --ref-members 2 --e2e-steps 1 2>&1 | tail -1 | grep -o "^{\"metric\": \"ensemble_member_steps_per_s\", \"value\": [0-9.]*\|launch_ms\": [0-9.]*\|failed[^\"]*"
done

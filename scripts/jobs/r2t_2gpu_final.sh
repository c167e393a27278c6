#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multigpu.py > gpurun_out/s8_check_multigpu.log 2>&1; echo "check rc=$?"; tail -6 gpurun_out/s8_check_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s8_bench_2gpu.json 2> gpurun_out/s8_bench_2gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s8_bench_2gpu.json

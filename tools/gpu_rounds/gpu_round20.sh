#!/bin/bash
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 tests/tp_worker.py > gpurun_out/tp2.log 2>&1; echo "tp2 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2_b32.log 2>&1

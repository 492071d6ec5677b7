#!/bin/bash
mkdir -p gpurun_out
python -m ctranslate2_b200.build > gpurun_out/build.log 2>&1
timeout 500 python bench.py --weights awq --steps 64 > gpurun_out/bench_awq_b32.log 2>&1
timeout 300 python bench.py --weights awq --batch 1 --steps 64 > gpurun_out/bench_awq_b1.log 2>&1
tail -c 1500 gpurun_out/bench_awq_b32.log; tail -c 600 gpurun_out/bench_awq_b1.log

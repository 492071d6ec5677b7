#!/bin/bash
timeout 600 python tools/gemm_bench.py gemm awq > gpurun_out/gemm_bench.log 2>&1
timeout 600 python bench.py --steps 64 --batch 32 --no-cpu-baseline > gpurun_out/bench_8b_b32.log 2> gpurun_out/bench_8b_b32.err; echo "bench8b_b32 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 64 --batch 1 --no-cpu-baseline > gpurun_out/bench_8b_b1.log 2> gpurun_out/bench_8b_b1.err; echo "bench8b_b1 rc=$?" >> gpurun_out/status.txt

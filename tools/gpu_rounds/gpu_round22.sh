#!/bin/bash
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
timeout 300 python tools/decode_once.py 32 32 > gpurun_out/base_b32.log 2>&1
timeout 300 python tools/decode_once.py 1 32 > gpurun_out/base_b1.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1

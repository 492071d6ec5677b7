#!/bin/bash
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
tail -n 1 gpurun_out/awq_b1.log gpurun_out/awq_b32.log

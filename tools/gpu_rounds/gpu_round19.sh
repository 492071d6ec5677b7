#!/bin/bash
mkdir -p gpurun_out
python -m ctranslate2_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_new_b32.log 2>&1
CT2B200_AWQ_DECODE_GLU=0 timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_mix_b32.log 2>&1
CT2B200_AWQ_DECODE_GLU=0 timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_mix_b1.log 2>&1
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_new_b1.log 2>&1
tail -n 1 gpurun_out/awq_*.log

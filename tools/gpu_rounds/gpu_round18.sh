#!/bin/bash
for g in dec1 dec32; do timeout 300 python tools/awq_probe.py $g > gpurun_out/awq_probe_$g.log 2>&1; done
CT2B200_AWQ_DECODE=1 timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_new_b1.log 2>&1
CT2B200_AWQ_DECODE=1 timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_new_b32.log 2>&1
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_old_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 32 > gpurun_out/base_b32.log 2>&1
timeout 300 python tools/decode_once.py 1 32 > gpurun_out/base_b1.log 2>&1

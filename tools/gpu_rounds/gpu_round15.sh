#!/bin/bash
for g in dec1 dec32 prefill; do timeout 300 python tools/awq_probe.py $g > gpurun_out/awq_probe_$g.log 2>&1; done
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python tools/decode_once.py 1 2 float16 8b awq_gemm > gpurun_out/awq_sanitizer.log 2>&1
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
CT2B200_AWQ_DECODE=0 timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_old_b1.log 2>&1

#!/bin/bash
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
CT2B200_ATTN_DECODE=persistent timeout 300 python tools/decode_once.py 32 32 > gpurun_out/persist_b32.log 2>&1
timeout 300 python tools/decode_once.py 64 16 > gpurun_out/auto_b64.log 2>&1
CT2B200_ATTN_DECODE=split timeout 300 python tools/decode_once.py 64 16 > gpurun_out/split_b64.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 330 --csv \
  --log-file gpurun_out/r01_launches_b32_v5.csv python tools/decode_once.py 32 2 > gpurun_out/ncu_list_b32.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.log 2>&1
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_reference.log 2>&1

#!/bin/bash
# AWQ baseline (current kernel), row-step experiment, ncu captures, default bench
( time timeout 600 python tools/decode_once.py 1 16 float16 8b awq_gemm ) > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
CT2B200_GEMM_ROWSTEP=1 timeout 300 python tools/decode_once.py 32 32 > gpurun_out/rowstep1_b32.log 2>&1
CT2B200_GEMM_ROWSTEP=1 timeout 300 python tools/decode_once.py 1 32 > gpurun_out/rowstep1_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 32 > gpurun_out/base_b32.log 2>&1
timeout 300 python tools/decode_once.py 1 32 > gpurun_out/base_b1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 330 --csv \
  --log-file gpurun_out/r01_launches_b32_v4.csv python tools/decode_once.py 32 2 > gpurun_out/ncu_list_b32.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 330 --csv \
  --log-file gpurun_out/r01_launches_b1_v4.csv python tools/decode_once.py 1 2 > gpurun_out/ncu_list_b1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"gemm_decode|attention_decode" -c 5 -f -o gpurun_out/r01_full_b32_v4 \
  python tools/decode_once.py 32 2 > gpurun_out/ncu_full_b32_v4.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:"gemm_prefill|attention_prefill" -c 12 --csv \
  --log-file gpurun_out/r01_launches_prefill_v4.csv python tools/decode_once.py 32 2 > gpurun_out/ncu_prefill.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.log 2>&1

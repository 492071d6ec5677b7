#!/bin/bash
mkdir -p gpurun_out
python -m ctranslate2_b200.build > gpurun_out/build.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"awq_decode" -c 4 -f \
  -o gpurun_out/r01_full_awq_b1 python tools/decode_once.py 1 2 float16 8b awq_gemm > gpurun_out/ncu_awq.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"gemm_prefill" -c 3 -f \
  -o gpurun_out/r01_full_prefill python tools/decode_once.py 32 2 > gpurun_out/ncu_prefill_full.log 2>&1
tail -n 2 gpurun_out/ncu_awq.log gpurun_out/ncu_prefill_full.log

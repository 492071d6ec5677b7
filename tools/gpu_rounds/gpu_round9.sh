#!/bin/bash
timeout 600 python tools/attn_sweep.py 32 0,1,2,3,4 > gpurun_out/attn_sweep_b32.log 2>&1
timeout 300 python tools/attn_sweep.py 1 0,4,8,16 > gpurun_out/attn_sweep_b1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"attention_decode_mma|gemm_tc" -c 10 -f -o gpurun_out/r01_full_b32_v3 \
  python tools/decode_once.py 32 2 > gpurun_out/ncu_full_b32_v3.log 2>&1
echo "ncu rc=$?" >> gpurun_out/status.txt

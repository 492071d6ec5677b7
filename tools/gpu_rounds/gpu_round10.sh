#!/bin/bash
timeout 600 python tools/ablate.py 32 32 > gpurun_out/ablate_b32.log 2>&1
timeout 600 python tools/ablate.py 1 32 > gpurun_out/ablate_b1.log 2>&1
for kb in 100 140; do
  CT2B200_GEMM_SMEM_KB=$kb timeout 300 python tools/decode_once.py 32 32 > gpurun_out/smem${kb}_b32.log 2>&1
  CT2B200_GEMM_SMEM_KB=$kb timeout 300 python tools/decode_once.py 1 32 > gpurun_out/smem${kb}_b1.log 2>&1
done
CT2B200_GEMM_DECODE=0 timeout 300 python tools/decode_once.py 32 32 > gpurun_out/nodecode_b32.log 2>&1
CT2B200_GEMM_DECODE=0 timeout 300 python tools/decode_once.py 1 32 > gpurun_out/nodecode_b1.log 2>&1

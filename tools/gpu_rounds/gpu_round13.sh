#!/bin/bash
timeout 600 python tools/ablate.py 32 32 > gpurun_out/ablate_b32.log 2>&1
timeout 600 python tools/ablate.py 1 32 > gpurun_out/ablate_b1.log 2>&1
CT2B200_ATTN_DECODE=split timeout 300 python tools/decode_once.py 32 32 > gpurun_out/attnsplit_b32.log 2>&1
CT2B200_ATTN_DECODE=split timeout 300 python tools/decode_once.py 1 32 > gpurun_out/attnsplit_b1.log 2>&1
CT2B200_GEMM_PREFILL=0 timeout 300 python tools/decode_once.py 32 8 > gpurun_out/noprefill_b32.log 2>&1
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/bench_b32.log 2>&1

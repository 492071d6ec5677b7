#!/bin/bash
timeout 600 python tools/gemm_bench.py gemm attn > gpurun_out/gemm_bench.log 2>&1
CT2B200_GEMM_WHOLE=1 timeout 300 python tools/gemm_bench.py gemm > gpurun_out/gemm_bench_whole.log 2>&1
CT2B200_PDL=0 timeout 300 python tools/gemm_bench.py gemm > gpurun_out/gemm_bench_nopdl.log 2>&1
timeout 300 python tools/gemm_bench.py awq > gpurun_out/awq_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 2 -o gpurun_out/prof_gemm python tools/gemm_bench.py gemm > gpurun_out/ncu_gemm.log 2>&1

#!/bin/bash
for g in dec1 dec32; do timeout 300 python tools/awq_probe.py $g > gpurun_out/awq_probe_$g.log 2>&1; done
timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 16 float16 8b awq_gemm > gpurun_out/awq_b32.log 2>&1
CT2B200_AWQ_DECODE=0 timeout 300 python tools/decode_once.py 1 16 float16 8b awq_gemm > gpurun_out/awq_old_b1.log 2>&1
timeout 300 python tools/decode_once.py 32 32 > gpurun_out/base_b32.log 2>&1
timeout 300 python tools/decode_once.py 1 32 > gpurun_out/base_b1.log 2>&1
CT2B200_ATTN_DECODE=persistent timeout 300 python tools/decode_once.py 32 32 > gpurun_out/persist_b32.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --tp --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tp2_b32.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --tp --batch 1 --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tp2_b1.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2_b32.log 2>&1

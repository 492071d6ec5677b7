#!/bin/bash
# profiling + bench round: launch lists, full captures of the two dominant kernels, B=32 / B=1 numbers
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --batch 32 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_b32.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b1.csv python bench.py --steps 2 --warmup 3 --batch 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_b1.log 2>&1
$NCU --set full --import-source on -k 'regex:gemm_tc_kernel|attention_decode_kernel' -c 5 -o gpurun_out/prof_b32 python bench.py --steps 1 --warmup 3 --batch 32 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_b32.log 2>&1
timeout 600 python bench.py --steps 128 --batch 32 --no-cpu-baseline > gpurun_out/bench_8b_b32.log 2> gpurun_out/bench_8b_b32.err; echo "bench8b_b32 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 128 --batch 1 --no-cpu-baseline > gpurun_out/bench_8b_b1.log 2> gpurun_out/bench_8b_b1.err; echo "bench8b_b1 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 8 --batch 32 > gpurun_out/bench_ref_b32.log 2> gpurun_out/bench_ref_b32.err; echo "bench_ref rc=$?" >> gpurun_out/status.txt

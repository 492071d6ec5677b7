#!/bin/bash
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 python bench.py --steps 64 --batch 32 --no-cpu-baseline > gpurun_out/bench_8b_b32.log 2> gpurun_out/bench_8b_b32.err; echo "bench8b_b32 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 64 --batch 1 --no-cpu-baseline > gpurun_out/bench_8b_b1.log 2> gpurun_out/bench_8b_b1.err; echo "bench8b_b1 rc=$?" >> gpurun_out/status.txt
CT2B200_PDL=0 timeout 600 python bench.py --steps 64 --batch 1 --no-cpu-baseline > gpurun_out/bench_8b_b1_nopdl.log 2> gpurun_out/bench_8b_b1_nopdl.err; echo "bench8b_b1_nopdl rc=$?" >> gpurun_out/status.txt
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --batch 32 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_b32.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b1.csv python bench.py --steps 2 --warmup 3 --batch 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_b1.log 2>&1

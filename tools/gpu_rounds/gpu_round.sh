#!/bin/bash
# One gpurun call: probes, GPU tests, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
python -m ctranslate2_b200.build > gpurun_out/build.log 2>&1
timeout 180 python tools/gpu_probe.py tc > gpurun_out/probe_tc.log 2>&1; echo "probe_tc rc=$?" >> gpurun_out/status.txt
timeout 180 python tools/gpu_probe.py mma > gpurun_out/probe_mma.log 2>&1; echo "probe_mma rc=$?" >> gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/status.txt
"$@"
cat gpurun_out/status.txt

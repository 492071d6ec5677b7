#!/bin/bash
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_worker.py > gpurun_out/tp2.log 2>&1; echo "tp2 rc=$?" >> gpurun_out/status.txt
timeout 600 python tools/ablate.py 32 32 > gpurun_out/ablate_b32.log 2>&1
timeout 600 python tools/ablate.py 1 32 > gpurun_out/ablate_b1.log 2>&1

#!/bin/bash
timeout 600 python tools/ablate.py 32 32 > gpurun_out/ablate_b32.log 2>&1
timeout 600 python tools/ablate.py 1 32 > gpurun_out/ablate_b1.log 2>&1

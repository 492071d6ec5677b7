#!/usr/bin/env python
"""Pipeline stamps of the AWQ decode kernel (CTA 0) at Llama-3-8B shapes, m = 32.  Needs the trace build:
  python -m ctranslate2_b200.build --variant awqtrace
  CT2B200_LIB=$PWD/ctranslate2_b200/libct2b200_awqtrace.so python tools/awq_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctranslate2_b200 import ops  # noqa: E402

dev = torch.device("cuda")


def awq_weight(n, k, g=128):
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), dtype=torch.int32, device=dev)
    sc = (torch.rand((k // g, n), device=dev) * 0.01 + 0.005).to(torch.float16)
    qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (k // g, n // 8), dtype=torch.int32, device=dev)
    return ops.AwqWeight(qw, sc, qz, ops.AWQ_GEMM, g)


m = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x4 = torch.randn((m, 4096), device=dev, dtype=torch.float16)
wg, wu = awq_weight(14336, 4096), awq_weight(14336, 4096)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(2):
    flush.fill_(i)                      # the second launch is the one to read: code and tensor maps warm, weights cold
    torch.cuda.synchronize()
    print("--- gate/up launch %d" % i, flush=True)
    ops.dense_awq_glu(x4, wg, wu)
    torch.cuda.synchronize()
wq = awq_weight(6144, 4096)
for i in range(2):
    flush.fill_(i)
    torch.cuda.synchronize()
    print("--- qkv launch %d" % i, flush=True)
    ops.dense_awq(x4, wq)
    torch.cuda.synchronize()

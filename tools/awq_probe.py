#!/usr/bin/env python
"""Op-by-op probe of the AWQ / float16 paths at Llama-3-8B shapes (sync after every call so a faulting kernel is named).
usage: python tools/awq_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctranslate2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
r = np.random.default_rng(0)


def step(name, fn, bytes_=0, reps=20):
    try:
        out = fn()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) * 1e3 / reps
        extra = " %.0f GB/s" % (bytes_ / us / 1e3) if bytes_ else ""
        print("OK   %-40s %.1f us%s" % (name, us, extra), flush=True)
        return out
    except Exception as e:  # noqa: BLE001
        print("FAIL %-40s %s" % (name, str(e)[:200]), flush=True)
        raise


def awq_weight(n, k, g=128):
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n // 8), dtype=torch.int32, device=dev)
    sc = (torch.rand((k // g, n), device=dev) * 0.01 + 0.005).to(torch.float16)
    qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (k // g, n // 8), dtype=torch.int32, device=dev)
    return ops.AwqWeight(qw, sc, qz, ops.AWQ_GEMM, g)


GROUP = sys.argv[1] if len(sys.argv) > 1 else "all"
shapes = {"qkv": (6144, 4096), "out": (4096, 4096), "down": (4096, 14336)}
for m in (1, 32):
    if GROUP not in ("all", "dec%d" % m):
        continue
    x4 = torch.randn((m, 4096), device=dev, dtype=torch.float16)
    x14 = torch.randn((m, 14336), device=dev, dtype=torch.float16)
    for name, (n, k) in shapes.items():
        w = awq_weight(n, k)
        x = x14 if k == 14336 else x4
        for env in ("1", "0"):
            os.environ["CT2B200_AWQ_DECODE"] = env
            step("dense_awq %s m=%d decode_kernel=%s" % (name, m, env), lambda: ops.dense_awq(x, w), n * k // 2)
        del w
    wg, wu = awq_weight(14336, 4096), awq_weight(14336, 4096)
    for env in ("1", "0"):
        os.environ["CT2B200_AWQ_DECODE"] = env
        step("dense_awq_glu m=%d decode_kernel=%s" % (m, env), lambda: ops.dense_awq_glu(x4, wg, wu), 14336 * 4096)
    del wg, wu
    lm = torch.randn((128256, 4096), device=dev, dtype=torch.float16) * 0.02
    step("gemm_f16 lm_head m=%d" % m, lambda: ops.Gemm()(x4, lm), 128256 * 4096 * 2, reps=5)
    del lm
if GROUP not in ("all", "prefill"):
    print("probe complete")
    sys.exit(0)
os.environ["CT2B200_AWQ_DECODE"] = "1"
x = torch.randn((1023, 4096), device=dev, dtype=torch.float16)
w = awq_weight(6144, 4096)
step("dense_awq qkv m=1023 (prefill)", lambda: ops.dense_awq(x, w), reps=3)
wg, wu = awq_weight(14336, 4096), awq_weight(14336, 4096)
step("dense_awq_glu m=1023 (prefill)", lambda: ops.dense_awq_glu(x, wg, wu), reps=3)
print("probe complete")

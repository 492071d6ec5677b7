#!/usr/bin/env python
"""One timed decode run of the 8B INT8 model (for ncu: bench_decode brackets the timed steps with
cudaProfilerStart/Stop).  usage: python tools/decode_once.py [batch] [steps] [compute_type] [model] [quant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ctranslate2_b200 as ct2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
CT = sys.argv[3] if len(sys.argv) > 3 else "int8_float16"
M = sys.argv[4] if len(sys.argv) > 4 else "8b"
P = 1024
QUANT = sys.argv[5] if len(sys.argv) > 5 else "int8_float16"
gen = ct2.Generator(bench.model_dir(M, QUANT), compute_type=CT, max_batch_size=B, max_length=P + 64 + 16)
pre, dec, n = gen.bench_decode(B, P, K, 3)
print("batch %d: prefill %.2f ms, decode %.3f ms/step, %d launches/step" % (B, pre, dec / K, n // K))

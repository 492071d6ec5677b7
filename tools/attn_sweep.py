#!/usr/bin/env python
"""Decode-attention slice-count sweep inside the real decode graph (CT2B200_ATTN_SPLITS is read at Generator open).
usage: python tools/attn_sweep.py [batch] [splits,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ctranslate2_b200 as ct2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SPL = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
P, K = 1024, 32
for s in SPL:
    if s:
        os.environ["CT2B200_ATTN_SPLITS"] = str(s)
    else:
        os.environ.pop("CT2B200_ATTN_SPLITS", None)
    gen = ct2.Generator(bench.model_dir("8b"), compute_type="int8_float16", max_batch_size=B, max_length=P + 64 + 16)
    os.environ["CT2B200_STEP_MASK"] = hex(0x3FF)
    _, dec, _ = gen.bench_decode(B, P, K, 3)
    os.environ["CT2B200_STEP_MASK"] = hex(0x3FF & ~4)
    _, dec2, _ = gen.bench_decode(B, P, K, 3)
    print("batch %d splits %s: step %.3f ms, attention class %.1f us/launch" % (B, s or "auto", dec / K, (dec - dec2) / K * 1e3 / 32))
    del gen

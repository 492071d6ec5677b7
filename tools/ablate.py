#!/usr/bin/env python
"""In-graph attribution of the decode step: times the CUDA-graph step with one kernel class removed at a time
(CT2B200_STEP_MASK), so per-class time includes launch gaps and PDL overlap exactly as in production.
usage: python tools/ablate.py [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ctranslate2_b200 as ct2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
P = 1024
NAMES = ["norm+quant(attn)", "qkv gemm", "attention", "quant(attn out)", "out gemm", "norm+quant(ffn)", "glu gemm",
         "quant(h)", "down gemm", "lm_head(+norm)"]
gen = ct2.Generator(bench.model_dir("8b"), compute_type="int8_float16", max_batch_size=B, max_length=P + K + 16)
FULL = 0x3FF


def run(mask):
    os.environ["CT2B200_STEP_MASK"] = hex(mask)
    _, dec, n = gen.bench_decode(B, P, K, 3)
    return dec / K, n // K


full, nl = run(FULL)
print("batch %d: full step %.3f ms (%d launches)" % (B, full, nl))
tot = 0.0
for i, name in enumerate(NAMES):
    ms, _ = run(FULL & ~(1 << i))
    print("  without %-18s %.3f ms  => class costs %.3f ms (%.1f us per launch)" % (name, ms, full - ms, (full - ms) * 1e3 / (32 if i < 9 else 1)))
    tot += full - ms
only_gemm, n2 = run(0b1101010010 & FULL)
print("  only the GEMMs: %.3f ms (%d launches); sum of class costs %.3f ms" % (only_gemm, n2, tot))
none, n3 = run(0)
print("  empty step (embedding + sampling): %.3f ms (%d launches)" % (none, n3))

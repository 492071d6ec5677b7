#!/usr/bin/env python
"""One timed run of the OPUS-MT-shaped translation step (for ncu: Translator.bench brackets the timed decoding steps with
cudaProfilerStart/Stop).  usage: python tools/translate_once.py [batch] [beam] [steps] [compute_type]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ctranslate2_b200.translator import Translator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
BEAM = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 64
CT = sys.argv[4] if len(sys.argv) > 4 else "int8_float16"
t = Translator(bench.seq2seq_model_dir(), compute_type=CT)
enc, dec, n = t.bench(B, 51, BEAM, K, 3)
print("batch %d beam %d: encoder %.3f ms, decode %.4f ms/step, %d launches/step" % (B, BEAM, enc, dec / K, n // K))

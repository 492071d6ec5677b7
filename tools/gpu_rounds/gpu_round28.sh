#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "scores or generate_matches_reference_tokens or ragged" --timeout 100 > gpurun_out/pytest_scores.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_scores.log
tail -3 gpurun_out/pytest_scores.log

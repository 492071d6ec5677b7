"""-m gpu: Generator (model dir -> load -> prefill -> greedy decode with CUDA graph) through the C-ABI with
HOST buffers, against (1) committed outputs of the unmodified reference on the reference-written tiny
model, (2) the oracle on a head_dim-128 synthetic model, (3) size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

import ctranslate2_b200 as ct2
from ctranslate2_b200.converters.synthetic import LlamaConfig, write_llama_model
from oracle import ct2_oracle as O
from oracle import refapi
from gpu_util import gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TINY = os.path.join(GOLDEN, "tiny_llama_int8")
FX = np.load(os.path.join(GOLDEN, "tiny_llama_int8_ref.npz"), allow_pickle=True)


@gpu
def rel_rms(a, ref):
    return float(np.sqrt(np.mean((a - ref).astype(np.float64) ** 2)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2)))


# Whole-model tolerances: the per-op tolerances of the reference (fp16 1e-2, bf16 4e-2) compound over the
# layers, and one flipped int8 rounding of an activation moves a whole row by 1/127 — so the end-to-end
# check is a relative RMS bound plus a looser max-abs bound; fp32 activations stay at 5e-4.  (On this 128-wide
# tiny model one int8 step is ~1 % of a row, so fp16 lands at ~2 % relative RMS; the head_dim-128 model below is tighter.)
@gpu
@pytest.mark.parametrize("compute_type,rms,mx", [("int8_float32", 2e-4, 5e-4), ("int8_float16", 4e-2, 1e-1),
                                                 ("int8_bfloat16", 1.5e-1, 4e-1)])
def test_tiny_model_logits_vs_reference(compute_type, rms, mx):
    g = ct2.Generator(TINY, compute_type=compute_type, max_batch_size=4, max_length=64)
    logits = g.forward_batch(FX["prompts"].tolist())
    ref = FX["logits"]
    assert rel_rms(logits, ref) <= rms, rel_rms(logits, ref)
    assert np.abs(logits - ref).max() <= mx * max(1.0, np.abs(ref).max()), np.abs(logits - ref).max()
    lp = g.forward_batch(FX["prompts"].tolist(), return_log_probs=True)
    np.testing.assert_allclose(np.exp(lp.astype(np.float64)).sum(-1), 1.0, atol=2e-2 if "32" not in compute_type else 1e-4)


@gpu
@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("impl", [1, 2])
def test_tiny_model_generate_matches_reference_tokens(graph, impl):
    """Greedy generate_batch: identical token ids to the unmodified reference (CPU, int8) — fp32 activations."""
    g = ct2.Generator(TINY, compute_type="int8_float32", max_batch_size=4, max_length=64, use_cuda_graph=graph,
                      gemm_impl=impl)
    prompts = FX["prompts"].tolist()
    res = g.generate_batch(prompts, max_length=12, min_length=0, end_token=[2])
    assert [r.sequences_ids[0] for r in res] == [list(x) for x in FX["generated"]]
    res = g.generate_batch(prompts, max_length=12, min_length=12, end_token=[2])
    assert [r.sequences_ids[0] for r in res] == FX["generated_min12"].tolist()
    # token strings come back through the vocabulary
    assert res[0].sequences[0][0] == "<t%d>" % res[0].sequences_ids[0][0]
    # a second call on the same generator (graph reuse, cache reuse) is identical
    res2 = g.generate_batch(prompts, max_length=12, min_length=12, end_token=[2])
    assert [r.sequences_ids[0] for r in res2] == FX["generated_min12"].tolist()


@gpu
def test_ragged_prompts_match_oracle():
    w = O.DecoderWeights.from_dir(TINY, "cuda")
    g = ct2.Generator(TINY, compute_type="int8_float32", max_batch_size=4, max_length=64)
    prompts = [[5, 9, 11, 40, 7], [8, 3, 77], [100, 23, 45, 67]]
    res = g.generate_batch(prompts, max_length=6, min_length=6, end_token=[2])
    for p, r in zip(prompts, res):       # each row alone through the oracle
        m = O.LlamaOracle(w)
        assert m.generate(np.array([p]), 6, 6, [2])[0] == r.sequences_ids[0]


@pytest.fixture(scope="module")
def d128_model(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("m128"))
    cfg = LlamaConfig(num_layers=2, num_heads=8, num_heads_kv=2, head_dim=128, ffn_dim=2048, vocab_size=2000,
                      rotary_scaling_type=2, rotary_scaling_factor=8.0, rotary_low_freq_factor=1.0,
                      rotary_high_freq_factor=4.0, original_max_position_embeddings=8192)
    write_llama_model(d, cfg, "int8_float16", seed=11, init_std=0.05)
    return d


# Whole-model tolerance with INT8 activations.  Every op matches the oracle to ~1e-6 with fp32 activations, EXCEPT
# where x*scale lands within an ulp of an int8 rounding boundary and two pipelines round it to different sides: that
# moves one activation by a full int8 step (1/127 of its row amax) and shifts the logits of that position by up to a
# few percent of their range.  It happens between the unmodified reference and the oracle too (measured on this very
# model: 9 of 200 single-token forwards differ by 0.07..0.17 on a logit range of 7, the other 191 agree to 2e-6).
# A flip in the input of the QKV projection changes that position's K/V and so reaches every later position of the row
# (reference vs oracle, one layer, 4x48 tokens: whole rows agree to 1e-7 or sit at 1e-2 of the logit range), so only a
# relative RMS bound is meaningful end to end.  Bit-level claims are made per op (test_gpu_ops.py); the smooth
# float16/bfloat16-weight arm of the same geometry is held to fp16 tolerance in test_awq_and_float_models_vs_oracle.
@gpu
@pytest.mark.parametrize("compute_type,rms", [("int8_float32", 8e-2), ("int8_float16", 8e-2)])
def test_d128_model_vs_oracle(d128_model, compute_type, rms):
    """Two layers: flips propagate through the KV cache, so the bound is a relative RMS (measured: 3.3% fp32, 4.9% fp16;
    reference vs oracle on the same model: 3.0%)."""
    w = O.DecoderWeights.from_dir(d128_model, "cuda")
    m = O.LlamaOracle(w)
    prompts = np.random.default_rng(3).integers(3, 2000, size=(2, 24))
    m.reset(2)
    ref = m.forward(prompts, 0)
    g = ct2.Generator(d128_model, compute_type=compute_type, max_batch_size=2, max_length=128)
    logits = g.forward_batch(prompts.tolist())
    assert rel_rms(logits, ref) <= rms, rel_rms(logits, ref)
    assert np.abs(logits - ref).max() <= 0.15 * np.abs(ref).max(), np.abs(logits - ref).max()
    ref_gen = m.generate(prompts, 16, 16, [2])
    res = g.generate_batch(prompts.tolist(), max_length=16, min_length=16, end_token=[2])
    got = [r.sequences_ids[0] for r in res]
    assert all(len(x) == 16 for x in got)
    if compute_type == "int8_float32":       # first generated token: same argmax unless it is a near-tie
        top2 = np.sort(ref[:, -1], axis=-1)[:, -2:]
        for b in range(2):
            if top2[b, 1] - top2[b, 0] > 0.3:
                assert got[b][0] == ref_gen[b][0]


@gpu
def test_decode_equals_prefill_property(d128_model):
    """Size-independent property (reference tests/model_test.cc:99-151): step-by-step decoding reproduces the
    full-sequence forward — here: greedy tokens are unchanged when the prompt is split differently between
    prefill and the decode loop (prompt forcing)."""
    g = ct2.Generator(d128_model, compute_type="int8_float16", max_batch_size=3, max_length=256)
    r = np.random.default_rng(9)
    base = r.integers(3, 2000, size=40).tolist()
    # row 0 alone: prefill 39 tokens.  In a ragged batch with a 5-token row: prefill 4, force 35 through the loop.
    alone = g.generate_batch([base], max_length=10, min_length=10, end_token=[2])[0].sequences_ids[0]
    mixed = g.generate_batch([base, base[:5]], max_length=10, min_length=10, end_token=[2])[0].sequences_ids[0]
    assert alone == mixed


@gpu
def test_errors():
    with pytest.raises(RuntimeError):
        ct2.Generator("/nonexistent/model")
    g = ct2.Generator(TINY, compute_type="int8_float16", max_batch_size=2, max_length=32)
    with pytest.raises(ValueError):
        g.generate_batch([[1, 2, 3]] * 3, max_length=4)          # batch > max_batch
    with pytest.raises(ValueError):
        g.generate_batch([[1, 2, 3]], max_length=64)             # exceeds the KV arena
    with pytest.raises(ValueError):
        g.generate_batch([[1, 2, 3]], max_length=4, beam_size=4)


@gpu
@pytest.mark.parametrize("quant", ["awq_gemm", "awq_gemv", "float16", "bfloat16"])
def test_awq_and_float_models_vs_oracle(tmp_path, quant):
    """AWQ-INT4 (both reference layouts) and float16/bfloat16-weight model dirs: logits vs the oracle, greedy
    tokens stable across prefill/decode splits."""
    d = str(tmp_path / quant)
    # ffn_dim != d_model on purpose: an AWQ_GEMM-packed weight is [K, N/8], so its first dimension is NOT the layer width
    cfg = LlamaConfig(num_layers=2, num_heads=8, num_heads_kv=2, head_dim=128, ffn_dim=1536, vocab_size=1000,
                      rotary_scaling_type=2, rotary_scaling_factor=8.0, rotary_low_freq_factor=1.0,
                      rotary_high_freq_factor=4.0, original_max_position_embeddings=64)
    write_llama_model(d, cfg, quant, seed=5, init_std=0.05)
    w = O.DecoderWeights.from_dir(d, "cuda")
    m = O.LlamaOracle(w)
    prompts = np.random.default_rng(4).integers(3, 1000, size=(2, 20))
    m.reset(2)
    ref = m.forward(prompts, 0)
    ctype = "bfloat16" if quant == "bfloat16" else "float16"
    g = ct2.Generator(d, compute_type=ctype, max_batch_size=2, max_length=128)
    logits = g.forward_batch(prompts.tolist())
    tol = 1.5e-1 if quant == "bfloat16" else 4e-2
    assert rel_rms(logits, ref) <= tol / 3, rel_rms(logits, ref)
    assert np.abs(logits - ref).max() <= tol * max(1.0, np.abs(ref).max()), np.abs(logits - ref).max()
    a = g.generate_batch(prompts.tolist(), max_length=8, min_length=8, end_token=[2])
    b = g.generate_batch(prompts.tolist(), max_length=8, min_length=8, end_token=[2])
    assert [r.sequences_ids[0] for r in a] == [r.sequences_ids[0] for r in b]      # deterministic


@gpu
def test_full_size_llama8b_properties():
    """BASELINE.json's full size (Llama-3-8B geometry, INT8, random weights): size-independent properties, since the
    oracle cannot run 8B in seconds — (1) the decode loop reproduces the one-shot forward: greedy tokens do not depend on
    how the prompt is split between the prompt pass and the forced decode steps; (2) generation is deterministic and
    independent of the batch neighbours; (3) every generated id is a valid vocabulary index."""
    import bench
    d = bench.model_dir("8b")
    g = ct2.Generator(d, compute_type="int8_float16", max_batch_size=4, max_length=512)
    r = np.random.default_rng(123)
    V = g.vocab_size
    base = r.integers(3, V, size=200).tolist()
    short = r.integers(3, V, size=9).tolist()
    alone = g.generate_batch([base], max_length=12, min_length=12, end_token=[1])[0].sequences_ids[0]
    # in a ragged batch the long row is forced through the decode loop from position 8 on
    mixed = g.generate_batch([base, short], max_length=12, min_length=12, end_token=[1])
    again = g.generate_batch([base, short], max_length=12, min_length=12, end_token=[1])
    assert [x.sequences_ids[0] for x in mixed] == [x.sequences_ids[0] for x in again]
    assert all(0 <= t < V for x in mixed for t in x.sequences_ids[0])
    assert len(alone) == 12 and len(mixed[0].sequences_ids[0]) == 12
    # fp16 activations + int8 rounding: the two schedules agree on the first tokens and mostly afterwards
    same = sum(int(a == b) for a, b in zip(alone, mixed[0].sequences_ids[0]))
    assert alone[0] == mixed[0].sequences_ids[0][0] or same >= 8, (alone, mixed[0].sequences_ids[0])


@gpu
@pytest.mark.parametrize("graph", [False, True])
def test_generate_scores_match_reference(graph):
    """return_scores: cumulative log-probabilities / length^penalty against the unmodified reference's
    GenerationResult.scores (tests/golden/tiny_llama_int8_scores.json, made by tools/make_golden.py --scores-only),
    incl. min_length (DisableTokens before LogSoftMax) and rows that stop on the end token."""
    import json
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_scores.json")))
    g = ct2.Generator(TINY, compute_type="int8_float32", max_batch_size=4, max_length=64, use_cuda_graph=graph)
    for c in fx["cases"]:
        res = g.generate_batch(fx["prompts"], max_length=c["max_length"], min_length=c["min_length"],
                               end_token=[c["end_id"]], return_scores=True, length_penalty=c["length_penalty"])
        assert [r.sequences_ids[0] for r in res] == c["tokens"]
        got = np.array([r.scores[0] for r in res], np.float32)
        np.testing.assert_allclose(got, np.array(c["scores"], np.float32), rtol=0, atol=2e-4 * (1 + np.abs(c["scores"]).max()))
    # without return_scores the result carries no scores and the tokens are the same
    res = g.generate_batch(fx["prompts"], max_length=12, min_length=12, end_token=[2])
    assert all(r.scores == [] for r in res)

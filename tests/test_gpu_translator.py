"""Encoder-decoder path on the GPU: ctranslate2_b200.Translator — encoder, cross-attention decoder and the
device-resident beam search — through the C-ABI, against (a) the committed outputs of the UNMODIFIED reference's Translator
(tests/golden/seq2seq_ref.json, tools/make_golden.py --seq2seq-only) and (b) the oracle run live.

Parity classes: float32 compute has no activation quantization, so tokens of every hypothesis must equal the reference's and
scores agree to 2e-4 (fp32 summation order); INT8 compute on a d=32 / d=64 model turns one rounding flip (an activation an ulp
from k + 0.5) into ~1e-2 of output, so int8 is pinned by the reference's golden sentence plus a majority agreement, exactly
as the oracle itself is pinned against the reference (tests/test_oracle_seq2seq.py)."""
import json
import os

import numpy as np
import pytest
import torch

import ctranslate2_b200 as ct2
from ctranslate2_b200 import ops
from ctranslate2_b200.translator import Translator
from oracle import ct2_oracle as O
from gpu_util import DEV, TDT, TOL, dev, gpu, round_through, to_np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = {"aren": os.path.join(GOLDEN, "aren-transliteration-i8"), "postnorm": os.path.join(GOLDEN, "tiny_seq2seq_postnorm")}
START, END = 1, 2          # <s>, </s> in both target vocabularies


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "seq2seq_ref.json")) as f:
        return json.load(f)


def _run(t, c):
    ids, lens, scores = t.translate_ids(c["sources"], beam_size=c["beam_size"], num_hypotheses=c["num_hypotheses"],
                                        max_decoding_length=c["max_length"], min_decoding_length=c["min_length"],
                                        length_penalty=c["length_penalty"], start_id=START, end_token=[END])
    hyps = [[ids[b, h, :lens[b, h]].tolist() for h in range(c["num_hypotheses"]) if lens[b, h] >= 0] for b in range(len(c["sources"]))]
    return hyps, [[float(scores[b, h]) for h in range(len(hyps[b]))] for b in range(len(hyps))]


# ---- op level -------------------------------------------------------------------------------------------------------
@gpu
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_layer_norm_matches_oracle(dtype):
    """ops::LayerNorm (layer_norm_gpu.cu:169-206) vs the oracle restatement pinned by the reference's LayerNorm golden."""
    r = np.random.default_rng(0)
    x = round_through(r.standard_normal((37, 96)) * 3 + 1, dtype)
    g, b = round_through(r.standard_normal(96), dtype), round_through(r.standard_normal(96), dtype)
    y = ops.LayerNorm(epsilon=1e-5)(dev(b, TDT[dtype]), dev(g, TDT[dtype]), dev(x, TDT[dtype]))
    np.testing.assert_allclose(to_np(y), O.layer_norm(x, g, b, 1e-5), atol=TOL[dtype] * 4, rtol=TOL[dtype])


@gpu
@pytest.mark.parametrize("round_before_cast", [True, False])
def test_layer_norm_quantize_is_the_two_ops(round_before_cast):
    """The fused LayerNorm + Quantize launch produces the bits of ops::Quantize applied to its own T output."""
    r = np.random.default_rng(1)
    x = dev(r.standard_normal((19, 128)).astype(np.float32) * 2)
    g, b = dev(r.standard_normal(128).astype(np.float32)), dev(r.standard_normal(128).astype(np.float32))
    y, q, s = ops.LayerNorm()(b, g, x, quantize=True, round_before_cast=round_before_cast)
    q2, s2 = ops.Quantize(round_before_cast)(y)
    assert torch.equal(q, q2) and torch.equal(s, s2)
    qo, so = O.quantize_rows(to_np(y), round_before_cast)
    assert np.array_equal(to_np(q), qo) and np.array_equal(to_np(s), so)


@gpu
@pytest.mark.parametrize("m,n,k", [(1, 43, 32), (7, 96, 32), (130, 70, 128), (64, 64, 64), (300, 33, 17)])
def test_gemm_f32_matches_fp64_truth(m, n, k):
    """primitives<CUDA>::gemm<float, float> (primitives.cu:485-505) + bias / activation / residual epilogue: true fp32."""
    r = np.random.default_rng(m + n + k)
    a, w = r.standard_normal((m, k)).astype(np.float32), r.standard_normal((n, k)).astype(np.float32)
    bias, res = r.standard_normal(n).astype(np.float32), r.standard_normal((m, n)).astype(np.float32)
    y = ops.Gemm()(dev(a), dev(w))
    np.testing.assert_allclose(to_np(y), a.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-5, atol=1e-5 * np.sqrt(k))
    y = ops.Gemm(activation_type=ops.ActivationType.ReLU)(dev(a), dev(w), dev(bias))
    np.testing.assert_allclose(to_np(y), np.maximum(a.astype(np.float64) @ w.astype(np.float64).T + bias, 0), rtol=1e-5,
                               atol=1e-5 * np.sqrt(k))
    y = ops.Gemm()(dev(a), dev(w), dev(bias), dev(res))
    np.testing.assert_allclose(to_np(y), a.astype(np.float64) @ w.astype(np.float64).T + bias + res, rtol=1e-5,
                               atol=1e-5 * np.sqrt(k))


# ---- engine level ----------------------------------------------------------------------------------------------------
@gpu
@pytest.mark.parametrize("name", ["aren", "postnorm"])
def test_float32_translations_equal_the_reference(fixture, name):
    t = Translator(MODELS[name], compute_type="float32")
    ref = fixture[name]["models"]["float32"]
    total = 0
    for c in ref["cases"]:
        hyps, scores = _run(t, c)
        assert hyps == c["hypotheses"], (c["sources"], c["beam_size"])
        for s, w in zip(scores, c["scores"]):
            np.testing.assert_allclose(s, w, atol=2e-4)
            total += len(s)
    assert total > 100
    t.close()


@gpu
@pytest.mark.parametrize("name", ["aren", "postnorm"])
def test_encoder_memory_matches_reference_and_oracle(fixture, name):
    t = Translator(MODELS[name], compute_type="float32")
    ref = fixture[name]["models"]["float32"]
    srcs = ref["encode_sources"]
    mem = t.encode(srcs)
    want = np.array(ref["memory"], np.float32)
    oracle = O.Seq2SeqOracle.from_dir(MODELS[name], compute_type="float32")
    S = max(len(r) for r in srcs)
    padded = np.zeros((len(srcs), S), np.int64)
    for b, r in enumerate(srcs):
        padded[b, :len(r)] = r
    mine = oracle.encode(padded, np.array([len(r) for r in srcs]))
    for b, r in enumerate(srcs):
        np.testing.assert_allclose(mem[b, :len(r)], want[b, :len(r)], atol=1e-4)
        np.testing.assert_allclose(mem[b, :len(r)], mine[b, :len(r)], atol=1e-4)
    t.close()


@gpu
@pytest.mark.parametrize("compute", ["int8", "int8_float16", "float16", "default"])
def test_golden_transliteration(compute):
    """tests/translator_test.cc:53-96 (ModelVariantTest): "آ ت ز م و ن" -> "a t z m o n", through the token-level API."""
    t = Translator(MODELS["aren"], compute_type=compute)
    res = t.translate_batch([["آ", "ت", "ز", "م", "و", "ن"]], beam_size=2, num_hypotheses=2, max_decoding_length=20,
                            return_scores=True)
    assert res[0].hypotheses[0] == ["a", "t", "z", "m", "o", "n"]
    assert len(res[0].hypotheses) == 2 and res[0].scores[0] >= res[0].scores[1]
    # the unmodified reference scores the two hypotheses -0.1553 / -0.2644 in int8 and -0.1554 / -0.2630 in float32
    assert abs(res[0].scores[0] - (-0.1553)) < 0.03 and abs(res[0].scores[1] - (-0.2644)) < 0.04
    assert res[0].hypotheses[1] == ["a", "t", "z", "u", "m", "o", "n"] or compute != "default"
    t.close()


@gpu
@pytest.mark.parametrize("name", ["aren", "postnorm"])
def test_int8_translations_agree_with_reference_statistically(fixture, name):
    """INT8 compute: most hypotheses are flip-free and equal the reference's; every first hypothesis stays close in score."""
    t = Translator(MODELS[name], compute_type="int8")
    ref = fixture[name]["models"]["int8"]
    same = total = close = 0
    for c in ref["cases"]:
        hyps, scores = _run(t, c)
        for b in range(len(hyps)):
            total += 1
            same += hyps[b][:1] == c["hypotheses"][b][:1]
            close += abs(scores[b][0] - c["scores"][b][0]) < 0.1
    # a d = 32 / 64 model turns one int8 rounding flip into a different best hypothesis now and then (the oracle itself is
    # pinned against the reference the same way): most sentences are flip-free, and the best scores stay close
    assert same / total >= 0.7, (same, total)
    assert close / total >= 0.6, (close, total)
    t.close()


@gpu
def test_cuda_graph_and_eager_steps_agree(fixture):
    ref = fixture["postnorm"]["models"]["float32"]["cases"]
    a = Translator(MODELS["postnorm"], compute_type="float32", use_cuda_graph=True)
    b = Translator(MODELS["postnorm"], compute_type="float32", use_cuda_graph=False)
    for c in ref[::5]:
        assert _run(a, c)[0] == _run(b, c)[0] == c["hypotheses"]
    n0 = ct2.kernel_launch_count()
    _run(a, ref[0])
    assert ct2.kernel_launch_count() > n0
    a.close()
    b.close()


@gpu
@pytest.mark.parametrize("beam,nh", [(4, 3), (7, 2), (10, 2)])
def test_beam_matches_oracle_on_long_batches(beam, nh):
    """Larger batch / longer decode than the fixture: 16 sources, 40 steps, vs the oracle live (float32).  Beam 4 and 7 take the
    one-pass scoring kernel with 8- and 16-entry per-thread lists, beam 10 the LogSoftMax + ops::TopK path."""
    t = Translator(MODELS["postnorm"], compute_type="float32")
    oracle = O.Seq2SeqOracle.from_dir(MODELS["postnorm"], compute_type="float32")
    rng = np.random.default_rng(5)
    srcs = [[int(x) for x in rng.integers(3, 120, size=int(rng.integers(3, 30)))] for _ in range(16)]
    c = dict(sources=srcs, beam_size=beam, num_hypotheses=nh, max_length=40, min_length=5, length_penalty=1.0)
    hyps, scores = _run(t, c)
    want = oracle.translate(srcs, beam_size=beam, num_hypotheses=nh, max_length=40, min_length=5, eos=END, bos=START)
    assert hyps == [[h[0] for h in w] for w in want]
    for s, w in zip(scores, want):
        np.testing.assert_allclose(s, [h[1] for h in w], atol=3e-4)
    t.close()


@gpu
def test_argument_errors():
    t = Translator(MODELS["aren"], compute_type="int8")
    with pytest.raises(ValueError):
        t.translate_batch([["a"]], target_prefix=[["b"]])
    with pytest.raises(ValueError):
        t.translate_batch([["a"]], beam_size=0)
    with pytest.raises(ValueError):
        t.translate_batch([["a"]], num_hypotheses=3, beam_size=2)
    with pytest.raises(ValueError):
        t.translate_batch([["a"]], sampling_topk=5)
    with pytest.raises(ValueError):
        t.translate_batch([["a"]], min_decoding_length=9, max_decoding_length=4)
    assert t.translate_batch([]) == []
    assert t.translate_batch([[]])[0].hypotheses == [[]]
    t.close()

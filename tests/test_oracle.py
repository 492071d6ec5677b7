"""The oracle (oracle/ct2_oracle.py) against (1) the golden vectors of the reference's own gtests,
(2) committed outputs of the unmodified reference, (3) the live reference when oracle/_ref is built."""
import json
import os

import numpy as np
import pytest

from oracle import ct2_oracle as O
from oracle import refapi

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GT = json.load(open(os.path.join(GOLDEN, "ref_gtest_vectors.json")))
RND = np.load(os.path.join(GOLDEN, "ref_ops_random.npz"))


def vec(test, name, dtype=np.float32, nth=0):
    items = [i for i in GT[test] if i["name"] == name]
    return np.array(items[nth]["values"], dtype=dtype).reshape(items[nth]["shape"])


# ---- the reference's own gtest goldens (tests/ops_test.cc, tests/layers_test.cc) ----

def test_gtest_gemm_int8():
    a, b = vec("GemmInt8", "a", np.int8), vec("GemmInt8", "b", np.int8)   # b is [K,N] there
    np.testing.assert_array_equal(O.gemm_s8(a, b.T.copy()), vec("GemmInt8", "expected", np.int32))


def test_gtest_quantize_int8():
    a = vec("QuantizeINT8", "a")
    q, s = O.quantize_rows(a, True)
    np.testing.assert_array_equal(q, vec("QuantizeINT8", "expected_qa", np.int8, 0))
    np.testing.assert_allclose(s, vec("QuantizeINT8", "expected_scale"), rtol=1e-6)
    q, _ = O.quantize_rows(a, False)   # legacy: no rounding before cast
    np.testing.assert_array_equal(q, vec("QuantizeINT8", "expected_qa", np.int8, 2))
    a = vec("QuantizeINT8ZeroRow", "a")
    q, s = O.quantize_rows(a, True)
    np.testing.assert_array_equal(q, vec("QuantizeINT8ZeroRow", "expected_qa", np.int8, 0))
    np.testing.assert_allclose(s, vec("QuantizeINT8ZeroRow", "expected_scale"), rtol=1e-6)


@pytest.mark.parametrize("test", ["TopK", "TopKVariableDepth"])
def test_gtest_topk(test):
    v, i = O.topk(vec(test, "input"), 3)
    np.testing.assert_array_equal(i, vec(test, "expected_indices", np.int32))
    np.testing.assert_allclose(v, vec(test, "expected_values"))
    v, i = O.topk(vec("TopKChangeK", "input"), 2)
    np.testing.assert_array_equal(i, vec("TopKChangeK", "expected_indices_k2", np.int32))


def test_gtest_softmax():
    np.testing.assert_allclose(O.softmax(vec("SoftMax", "x")), vec("SoftMax", "expected"), atol=1e-5)
    np.testing.assert_allclose(O.softmax(vec("LogSoftMax", "x"), log=True), vec("LogSoftMax", "expected"), atol=1e-5)
    y = O.softmax(vec("MaskedSoftMax", "x"), vec("MaskedSoftMax", "lengths", np.int32))
    np.testing.assert_allclose(y, vec("MaskedSoftMax", "expected"), atol=1e-5)


def test_gtest_rms_norm():
    y = O.rms_norm(vec("RMSNorm", "x"), vec("RMSNorm", "gamma"), 1e-6)
    np.testing.assert_allclose(y, vec("RMSNorm", "expected"), atol=1e-5)


@pytest.mark.parametrize("test,act", [("Swish", O.ACT_SWISH), ("ReLU", O.ACT_RELU), ("GELU", O.ACT_GELU),
                                      ("GELUTanh", O.ACT_GELU_TANH), ("GELUSigmoid", O.ACT_GELU_SIGMOID)])
def test_gtest_activations(test, act):
    np.testing.assert_allclose(O.activation(vec(test, "input"), act), vec(test, "expected"), atol=1e-5)


def static_vec(name):
    it = next(i for i in GT["_static"] if i["name"] == name)
    return np.array(it["values"], np.float32).reshape(it["shape"])


def test_gtest_gemm_float_bias_residual_gelu():
    """tests/ops_test.cc:516-681 — the float Gemm arm with beta * C, bias, residual and the GELU epilogue (act after both)."""
    a, b, y = static_vec("gemm_a"), static_vec("gemm_b"), static_vec("gemm_y")
    tol = dict(atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(O.gemm_float(a, b, y, 1.0, 1.0), vec("Gemm", "expected"), **tol)
    np.testing.assert_allclose(O.gemm_float(a.T.copy(), b, y, 1.0, 1.0, trans_a=True), vec("Gemm", "expected"), **tol)
    np.testing.assert_allclose(O.gemm_float(a, b.T.copy(), y, 1.0, 1.0, trans_b=True), vec("Gemm", "expected"), **tol)
    np.testing.assert_allclose(O.gemm_float(a, b, y, 1.0, 1.0, bias=vec("GemmBias", "bias")), vec("GemmBias", "expected"), **tol)
    res = vec("GemmResidual", "residual")
    np.testing.assert_allclose(O.gemm_float(a, b, y, 1.0, 1.0, bias=vec("GemmResidual", "bias"), residual=res),
                               vec("GemmResidual", "expected", nth=0), **tol)
    np.testing.assert_allclose(O.gemm_float(a, b, y, 1.0, 1.0, residual=res), vec("GemmResidual", "expected", nth=1), **tol)
    bias, res = vec("GemmGELU", "bias"), vec("GemmGELU", "residual")
    np.testing.assert_allclose(O.gemm_float(a, b, bias=bias, residual=res, act=O.ACT_GELU),
                               vec("GemmGELU", "expected", nth=0), **tol)
    np.testing.assert_allclose(O.gemm_float(a, b, bias=bias, act=O.ACT_GELU), vec("GemmGELU", "expected", nth=1), **tol)
    np.testing.assert_allclose(O.gemm_float(a, b, act=O.ACT_GELU), vec("GemmGELU", "expected", nth=2), **tol)


def test_gtest_layer_norm():
    """tests/ops_test.cc:880-895 — LayerNorm with gamma and beta (the encoder-decoder restatement uses it)."""
    y = O.layer_norm(vec("LayerNorm", "x"), vec("LayerNorm", "gamma"), vec("LayerNorm", "beta"), 1e-5)
    np.testing.assert_allclose(y, vec("LayerNorm", "expected"), atol=1e-5)


def test_gtest_bias_add():
    """tests/ops_test.cc:1398-1432 — BiasAdd with the GELU epilogue, last axis and axis -2."""
    value, bias = static_vec("bias_value"), static_vec("bias_bias")
    np.testing.assert_allclose(O.bias_add(value, bias, O.ACT_GELU), vec("BiasAddGELU", "expected"), atol=1e-5)
    np.testing.assert_allclose(O.bias_add(value, bias, O.ACT_GELU, axis=-2), vec("BiasAddAxisGELU", "expected"), atol=1e-5)


def test_gtest_rotary_embedding():
    x, exp = vec("RotaryEmbedding", "input"), vec("RotaryEmbedding", "expected")
    # default RotaryEmbeddings(): dim 0 (= depth), interleave, base 10000, applied at offset 2
    sin, cos = O.rotary_tables(4, 6, 10000.0, interleave=True)
    np.testing.assert_allclose(O.rotary(x, sin[2:4], cos[2:4], True), exp, atol=1e-5)

    def permute(t):
        return t.reshape(8, 2, 3, 2).transpose(0, 1, 3, 2).reshape(2, 4, 2, 6)
    sin, cos = O.rotary_tables(4, 6, 10000.0, interleave=False)
    np.testing.assert_allclose(O.rotary(permute(x), sin[2:4], cos[2:4], False), permute(exp), atol=1e-5)


def test_gtest_gather():
    for t in ("GatherData1D", "GatherData2D", "GatherData3D"):
        np.testing.assert_array_equal(O.gather_rows(vec(t, "data"), vec(t, "ids", np.int32)), vec(t, "expected"))


# ---- committed outputs of the unmodified reference on seeded random inputs ----

def test_ref_random_quantize_gemm_dequantize():
    q, s = O.quantize_rows(RND["q_x"])
    np.testing.assert_array_equal(q, RND["q_q"])
    np.testing.assert_array_equal(s, RND["q_s"])
    np.testing.assert_array_equal(O.gemm_s8(RND["g_a"], RND["g_b"]), RND["g_c"])
    for act in (-1, 0, 1, 2, 3, 4, 5, 6):
        y = O.dequantize_gemm_output(RND["g_c"], RND["dq_sa"], RND["dq_sb"], RND["dq_bias"], act, "cpu")
        # GELU/tanh use the reference's vectorised erf/tanh approximations on CPU (avx_mathfun): ~2e-5 abs
        tol = 3e-5 if act in (1, 3, 5) else 2e-6
        np.testing.assert_allclose(y, RND["dq_y_act%d" % act], rtol=tol, atol=tol)
        y = O.dequantize_gemm_output(RND["g_c"], RND["dq_sa"], RND["dq_sb"], RND["dq_bias"], act, "cuda")
        np.testing.assert_allclose(y, RND["dq_y_act%d" % act], rtol=3e-5, atol=3e-5)


def test_ref_random_norm_rotary_softmax_topk_gather():
    np.testing.assert_allclose(O.rms_norm(RND["q_x"], RND["rn_gamma"], 1e-5), RND["rn_y"], rtol=1e-5, atol=1e-6)
    for key, inter in (("ro_y_interleave", True), ("ro_y_half", False)):
        np.testing.assert_allclose(O.rotary(RND["ro_x"], RND["ro_sin"], RND["ro_cos"], inter), RND[key], atol=1e-6)
    np.testing.assert_allclose(O.softmax(RND["sm_x"]), RND["sm_y"], atol=1e-6)
    np.testing.assert_allclose(O.softmax(RND["sm_x"], RND["sm_len"]), RND["sm_y_len"], atol=1e-6)
    np.testing.assert_allclose(O.softmax(RND["sm_x"], None, True), RND["sm_logy"], atol=1e-5)
    for k in (1, 4):
        v, i = O.topk(RND["tk_x"], k)
        np.testing.assert_array_equal(v, RND["tk_v%d" % k])
        ref_i = RND["tk_i%d" % k]
        if k == 1:
            np.testing.assert_array_equal(i, ref_i)
        else:
            # k>1 on CPU is std::partial_sort with a strict `>` comparator (topk_cpu.cc:38-41): the order
            # of exactly-tied values is unspecified there (row 1 comes back as [500, 17]); ours is
            # lowest-index-first.  Rows without ties must agree exactly, tied rows as sets.
            np.testing.assert_array_equal(np.delete(i, 1, 0), np.delete(ref_i, 1, 0))
            assert sorted(i[1]) == sorted(ref_i[1])
    assert RND["tk_i1"][1, 0] == 17          # exact tie 17 vs 500: lowest index wins in the reference
    np.testing.assert_array_equal(O.gather_rows(RND["ga_d"], RND["ga_i"]), RND["ga_y"])


def test_tiny_llama_against_reference_fixture():
    fx = np.load(os.path.join(GOLDEN, "tiny_llama_int8_ref.npz"), allow_pickle=True)
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), flavor="cpu")
    m = O.LlamaOracle(w)
    prompts = fx["prompts"].astype(np.int64)
    m.reset(prompts.shape[0])
    logits = m.forward(prompts, 0)
    np.testing.assert_allclose(logits, fx["logits"], atol=2e-5)
    gen = m.generate(prompts, 12, 0, [2])
    assert gen == [list(g) for g in fx["generated"]]
    gen = m.generate(prompts, 12, 12, [2])
    assert gen == fx["generated_min12"].tolist()
    # step-by-step decode == full forward (reference tests/model_test.cc:99-151 DecoderIterativeSequence)
    m.reset(prompts.shape[0])
    steps = [m.forward(prompts[:, t:t + 1], t)[:, 0] for t in range(prompts.shape[1])]
    np.testing.assert_allclose(np.stack(steps, 1), logits, atol=2e-5)


def test_awq_layouts_agree():
    r = np.random.default_rng(3)
    K, N, G = 256, 64, 128
    w_int = r.integers(0, 16, size=(K, N))
    z_int = r.integers(0, 16, size=(K // G, N))
    scales = (r.uniform(0.005, 0.02, size=(K // G, N))).astype(np.float16)
    qw, qz = O.awq_pack_gemm(w_int, z_int)
    np.testing.assert_array_equal(O.awq_unpack_gemm(qw), w_int)
    x = r.standard_normal((3, K)).astype(np.float32)
    y_gemm = O.awq_gemm(x, qw, scales, qz)
    qw2, qz2, sc2 = O.awq_pack_gemv(w_int.T.copy(), z_int.T.copy(), scales.T.copy(), G)
    y_gemv = O.awq_gemv(x, qw2, sc2, qz2, G)
    np.testing.assert_allclose(y_gemm, y_gemv, rtol=1e-5, atol=1e-5)
    deq = (w_int - np.repeat(z_int, G, 0)) * np.repeat(scales.astype(np.float32), G, 0)
    np.testing.assert_allclose(y_gemm, x @ deq, rtol=1e-5, atol=1e-5)


# ---- live cross-check against the compiled reference, when it is present ----

@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref not built")
def test_live_reference_ops():
    r = np.random.default_rng(11)
    x = (r.standard_normal((9, 200)) * 2).astype(np.float32)
    q, s = O.quantize_rows(x)
    rq, rs = refapi.quantize(x)
    np.testing.assert_array_equal(q, rq)
    np.testing.assert_array_equal(s, rs)
    b = r.integers(-127, 128, size=(40, 200), dtype=np.int8)
    np.testing.assert_array_equal(O.gemm_s8(q, b), refapi.gemm_s8(q, b))
    v, i = O.topk(x, 5)
    rv, ri = refapi.topk(x, 5)
    np.testing.assert_array_equal(i, ri)


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref not built")
def test_live_reference_synthetic_writer_roundtrip(tmp_path):
    """A model dir written by OUR writer loads in the unmodified reference and the oracle agrees."""
    from ctranslate2_b200.converters.synthetic import LlamaConfig, write_llama_model
    cfg = LlamaConfig(num_layers=2, num_heads=4, num_heads_kv=2, head_dim=32, ffn_dim=192, vocab_size=150)
    mdir = str(tmp_path / "m")
    write_llama_model(mdir, cfg, "int8", seed=5)
    g = refapi.RefGenerator(mdir, "int8", 2)
    prompts = np.random.default_rng(0).integers(3, 150, size=(2, 6), dtype=np.int32)
    ref_logits = g.forward(prompts)
    ref_gen = g.generate(prompts, 10, 10, 2)
    g.close()
    m = O.LlamaOracle(O.DecoderWeights.from_dir(mdir, "cpu"))
    m.reset(2)
    np.testing.assert_allclose(m.forward(prompts.astype(np.int64), 0), ref_logits, atol=2e-5)
    assert m.generate(prompts.astype(np.int64), 10, 10, [2]) == ref_gen


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("stored,requested", [("float32", "int8"), ("float16", "int8"), ("int8", "float32")])
def test_live_reference_on_load_compute_type_conversion(tmp_path, stored, requested):
    """Model::set_compute_type / ensure_dtype (model.cc:178-234, 304-369): a float model requested as int8 is quantized at
    load, an int8 model requested as float32 is dequantized — the oracle's restatement against the unmodified reference."""
    from ctranslate2_b200.converters.synthetic import LlamaConfig, write_llama_model
    cfg = LlamaConfig(num_layers=2, num_heads=4, num_heads_kv=2, head_dim=32, ffn_dim=192, vocab_size=150)
    mdir = str(tmp_path / "m")
    write_llama_model(mdir, cfg, stored, seed=6, init_std=0.05)
    g = refapi.RefGenerator(mdir, requested, 2)
    prompts = np.random.default_rng(1).integers(3, 150, size=(2, 7), dtype=np.int32)
    ref_logits = g.forward(prompts)
    g.close()
    w = O.ensure_compute_type(O.DecoderWeights.from_dir(mdir, "cpu"), "int8" if requested == "int8" else "float")
    if stored == "float16":      # non-weight variables (norms) are converted to float32 on the CPU (is_convertible branch)
        w.v.update({k: a.astype(np.float32) for k, a in w.v.items() if a.dtype == np.float16})
    m = O.LlamaOracle(w)
    m.reset(2)
    np.testing.assert_allclose(m.forward(prompts.astype(np.int64), 0), ref_logits, atol=5e-5)


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(), dict(scaling_type=0, scaling_factor=4.0),
                                dict(scaling_type=2, scaling_factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                     original_max_position_embeddings=8192)])
def test_live_reference_rotary_tables(kw):
    """RotaryEmbeddings::initialize incl. Linear and Llama-3 frequency scaling: oracle tables == the reference's."""
    for dim in (64, 128):
        rs, rc = refapi.rotary_tables(300, dim, 500000.0, False, **kw)
        s, c = O.rotary_tables(300, dim, 500000.0, False, **kw)
        # angles reach 300 rad, where one float ulp is 3e-5: libm pow/sin vs numpy differ by an ulp of the angle
        np.testing.assert_allclose(s, rs, atol=1e-4)
        np.testing.assert_allclose(c, rc, atol=1e-4)
        assert np.abs(s - rs).mean() < 2e-6


def test_tensor_parallel_partition_restatement():
    """Tensor-parallel partition (model.cc:662-743) restated in the oracle: column-parallel shards concatenate to the
    unsharded output bit for bit; row-parallel INT8 partials (global per-row amax) sum to the unsharded output up to
    fp32 rounding, because the int32 accumulators of the K slices add up exactly."""
    r = np.random.default_rng(0)
    H, Hkv, D, d, F = 8, 4, 16, 128, 256
    x = r.standard_normal((5, d)).astype(np.float32)
    wq, ws = O.quantize_weight((r.standard_normal(((H + 2 * Hkv) * D, d)) * 0.05).astype(np.float32))
    full = O.dense_int8(x, wq, ws)
    for world in (2, 4):
        parts = []
        for rank in range(world):
            rows = O.tp_qkv_rows(H, Hkv, D, rank, world)
            parts.append((rows, O.dense_int8(x, wq[rows], ws[rows])))
        # every rank holds whole heads: H/world query heads followed by Hkv/world key and value heads
        assert all(p[1].shape[1] == (H + 2 * Hkv) * D // world for p in parts)
        recon = np.zeros_like(full)
        for rows, y in parts:
            recon[:, rows] = y
        np.testing.assert_array_equal(recon, full)
        wd, sd = O.quantize_weight((r.standard_normal((d, F)) * 0.05).astype(np.float32))
        h = r.standard_normal((5, F)).astype(np.float32)
        res = r.standard_normal((5, d)).astype(np.float32)
        np.testing.assert_allclose(O.dense_int8_row_parallel(h, wd, sd, world, res), O.dense_int8(h, wd, sd, None, -1, res),
                                   rtol=0, atol=2e-6 * float(np.abs(res).max() + 1))
        # the exact statement behind it: sliced int32 GEMMs add up to the full GEMM
        hq, _ = O.quantize_rows(h)
        acc = sum(O.gemm_s8(hq[:, slice(*O.tp_shard_rows(F, k, world))], wd[:, slice(*O.tp_shard_rows(F, k, world))])
                  for k in range(world))
        np.testing.assert_array_equal(acc, O.gemm_s8(hq, wd))


def test_generate_scores_match_reference_fixture():
    """GenerationOptions::return_scores (decoding.cc:875-923, 189-203): the oracle's cumulative log-probabilities,
    length-normalised, against the unmodified reference's GenerationResult.scores on the committed tiny model —
    incl. rows that stop on the end token (its log-probability is part of the score, the token is not returned)."""
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_scores.json")))
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), "cpu")
    m = O.LlamaOracle(w)
    prompts = np.array(fx["prompts"])
    for c in fx["cases"]:
        toks, scores = m.generate(prompts, c["max_length"], c["min_length"], [c["end_id"]], return_scores=True,
                                  length_penalty=c["length_penalty"])
        assert toks == c["tokens"]
        np.testing.assert_allclose(scores, np.array(c["scores"], np.float32), rtol=0, atol=2e-5)


def test_ragged_prompts_match_reference_fixture():
    """Prompts of different lengths in one batch (language_model.cc:217-238): every row of the reference's result equals
    the row decoded ALONE by the oracle — tokens and scores — which is the semantics the engine implements (forced prompt
    tokens, per-row min_length).  Exception, recorded on purpose: when the SHORTEST prompt is a single token the reference
    never clears return_prefix, so longer rows come back with their forced prompt tokens in front, counted against
    max_length and in the score; the engine does not reproduce that."""
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_ragged.json")))
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), "cpu")
    regular = quirk = 0
    for c in fx["cases"]:
        lens = [len(p) for p in c["prompts"]]
        for b, p in enumerate(c["prompts"]):
            toks, scores = O.LlamaOracle(w).generate(np.array([p]), c["max_length"], c["min_length"], [c["end_id"]],
                                                     return_scores=True)
            if min(lens) >= 2 or len(p) == 1 or max(lens) == 1:
                assert toks[0] == c["tokens"][b], (c["prompts"], b)
                np.testing.assert_allclose(scores[0], c["scores"][b], rtol=0, atol=2e-5)
                regular += 1
            else:       # the reference returns the forced prompt tokens first
                assert c["tokens"][b][:len(p) - 1] == p[1:][:c["max_length"]]
                quirk += 1
    assert regular >= 30 and quirk >= 4


def test_score_batch_matches_reference_fixture():
    """Generator::score_batch (src/scoring.cc:6-66): log-probability of every token given its prefix — ragged batch, sequences
    too short to score, ScoringOptions::offset."""
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_score_batch.json")))
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), "cpu")
    m = O.LlamaOracle(w)
    for c in fx["cases"]:
        got = m.score(c["sequences"], c["offset"])
        assert [len(x) for x in got] == [len(x) for x in c["log_probs"]]
        for a, b in zip(got, c["log_probs"]):
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


def test_logits_processors_match_reference_fixture():
    """GenerationOptions::repetition_penalty / no_repeat_ngram_size / disable_unk / suppress_sequences through the greedy loop
    (decoding.cc:845-850, decoding_utils.cc): tokens identical, scores to fp32 round-off, incl. end tokens and min_length."""
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_processors.json")))
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), "cpu")
    m = O.LlamaOracle(w)
    prompts = np.array(fx["prompts"])
    changed = 0
    plain = {}
    for c in fx["cases"]:
        opt = dict(c["options"])
        if opt.pop("disable_unk", False):
            opt["disable_ids"] = [0]                   # <t0> is the unknown token of the tiny model
        key = (c["max_length"], c["min_length"], c["end_id"])
        if key not in plain:
            plain[key] = m.generate(prompts, c["max_length"], c["min_length"], [c["end_id"]])
        toks, scores = m.generate(prompts, c["max_length"], c["min_length"], [c["end_id"]], return_scores=True, **opt)
        assert toks == c["tokens"], c["options"]
        np.testing.assert_allclose(scores, c["scores"], rtol=0, atol=2e-5)
        changed += int(toks != plain[key])
    assert changed >= len(fx["cases"]) // 2          # the options really alter what is generated


def test_gtest_penalize_previous_tokens():
    """tests/primitives_test.cc:33-52 (PenalizePreviousTokens): penalty 1.2; row 0 saw token 2 twice (penalised once), row 1
    saw tokens 1 and 2."""
    scores = np.array([[0.6, 0.2, -1.2, 0.1], [0.3, 0.5, -1.3, 0.2]], np.float32)
    expected = scores.copy()
    expected[0, 2] *= np.float32(1.2)
    expected[1, 1] /= np.float32(1.2)
    expected[1, 2] *= np.float32(1.2)
    for row, prev in zip(scores, ([2, 2], [1, 2])):
        O.apply_logits_processors(row, prev, repetition_penalty=1.2)
    np.testing.assert_array_equal(scores, expected)


def test_logits_processor_rules():
    lowest = np.finfo(np.float32).min
    base = np.array([1.0, -2.0, 3.0, 0.5, -0.5], np.float32)
    x = base.copy()
    O.apply_logits_processors(x, [2, 1, 2], repetition_penalty=2.0)          # a token seen twice is penalised once
    np.testing.assert_array_equal(x, np.array([1.0, -4.0, 1.5, 0.5, -0.5], np.float32))
    x = base.copy()
    O.apply_logits_processors(x, [3, 4, 0, 3], no_repeat_ngram_size=2)         # "3 4" happened: after 3, token 4 is banned
    assert x[4] == lowest and (x[:4] == base[:4]).all()
    x = base.copy()
    O.apply_logits_processors(x, [3, 4], no_repeat_ngram_size=3)               # shorter than the n-gram: nothing to ban
    np.testing.assert_array_equal(x, base)
    x = base.copy()
    O.apply_logits_processors(x, [], suppress_sequences=[[1], [0, 2]])          # step 0: single tokens only
    assert x[1] == lowest and x[2] == base[2]
    x = base.copy()
    O.apply_logits_processors(x, [4, 0], suppress_sequences=[[1], [0, 2]])
    assert x[1] == lowest and x[2] == lowest


def test_beam_search_matches_reference_fixture():
    """BeamSearch::search (decoding.cc:425-720) restated in the oracle (the next row): hypotheses and scores of
    the unmodified reference's Generator with beam_size 2 / 4, several end tokens (rows finishing at step 0, mid-sequence,
    never), min_length, num_hypotheses, patience, with and without length penalty."""
    fx = json.load(open(os.path.join(GOLDEN, "tiny_llama_int8_scores.json")))
    w = O.DecoderWeights.from_dir(os.path.join(GOLDEN, "tiny_llama_int8"), "cpu")
    m = O.LlamaOracle(w)
    prompts = np.array(fx["prompts"])
    for c in fx["beam_cases"]:
        got = m.generate_beam(prompts, c["beam_size"], c["max_length"], c["min_length"], [c["end_id"]], c["length_penalty"],
                              c["num_hypotheses"], c["patience"])
        for row_got, row_ref in zip(got, c["hypotheses"]):
            assert [t for t, _ in row_got] == [t for t, _ in row_ref], c
            np.testing.assert_allclose([s for _, s in row_got], [s for _, s in row_ref], rtol=0, atol=5e-5)

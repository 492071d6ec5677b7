"""Encoder-decoder path on the CPU side: the oracle's Seq2SeqOracle against the COMMITTED outputs of the
unmodified reference's Translator (tests/golden/seq2seq_ref.json, written by tools/make_golden.py --seq2seq-only), on the
reference's own golden model (tests/translator_test.cc:53-96) and on a post-norm / Swish / start-from-zero model (the OPUS-MT
recipe in small); plus the host-only model parser of the engine.  No GPU, no /root/reference."""
import json
import os

import numpy as np
import pytest

from oracle import ct2_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = {"aren": os.path.join(GOLDEN, "aren-transliteration-i8"), "postnorm": os.path.join(GOLDEN, "tiny_seq2seq_postnorm")}


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "seq2seq_ref.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["aren", "postnorm"])
def test_oracle_float32_matches_reference_fixture(fixture, name):
    """float32 compute has no activation quantization: every token of every hypothesis must agree, scores to 1e-4."""
    oracle = O.Seq2SeqOracle.from_dir(MODELS[name], compute_type="float32")
    ref = fixture[name]["models"]["float32"]
    hyps = 0
    for c in ref["cases"]:
        got = oracle.translate(c["sources"], beam_size=c["beam_size"], num_hypotheses=c["num_hypotheses"],
                               max_length=c["max_length"], min_length=c["min_length"], length_penalty=c["length_penalty"])
        for g, toks, scores in zip(got, c["hypotheses"], c["scores"]):
            assert [h[0] for h in g] == toks
            np.testing.assert_allclose([h[1] for h in g], scores, atol=1e-4)
            hyps += len(g)
    assert hyps > 100
    srcs = ref["encode_sources"]
    S = max(len(r) for r in srcs)
    padded = np.zeros((len(srcs), S), np.int64)
    for b, r in enumerate(srcs):
        padded[b, :len(r)] = r
    mine = oracle.encode(padded, np.array([len(r) for r in srcs]))
    want = np.array(ref["memory"], np.float32)
    for b, r in enumerate(srcs):
        np.testing.assert_allclose(mine[b, :len(r)], want[b, :len(r)], atol=3e-5)


def test_golden_transliteration_from_the_committed_model():
    """tests/translator_test.cc:53-96 through the oracle on the committed copy of the reference's model."""
    mdir = MODELS["aren"]
    src = [l.rstrip("\n") for l in open(os.path.join(mdir, "source_vocabulary.txt"), encoding="utf-8")]
    tgt = [l.rstrip("\n") for l in open(os.path.join(mdir, "target_vocabulary.txt"), encoding="utf-8")]
    ids = [src.index(w) for w in ["آ", "ت", "ز", "م", "و", "ن"]]
    hyp = O.Seq2SeqOracle.from_dir(mdir).translate([ids], beam_size=2, max_length=20)[0][0][0]
    assert [tgt[i] for i in hyp] == ["a", "t", "z", "m", "o", "n"]


def test_translator_summary_parses_both_models():
    from ctranslate2_b200.translator import translator_summary
    a = translator_summary(MODELS["aren"])
    assert a["spec"] == "TransformerBase" and a["binary_version"] == 2 and a["encoder_layers"] == 6 and a["num_heads"] == 8
    assert a["d_model"] == 32 and a["weights"] == "int8" and a["pre_norm"] is True and a["round_before_cast"] is False
    assert abs(a["embeddings_scale"] - 32 ** 0.5) < 1e-6
    p = translator_summary(MODELS["postnorm"])
    assert p["spec"] == "TransformerSpec" and p["binary_version"] == 6 and p["pre_norm"] is False and p["activation"] == 2
    assert p["decoder_layers"] == 2 and p["head_dim"] == 16 and p["target_vocab"] == 96 and p["round_before_cast"] is True
    with pytest.raises(ValueError):
        translator_summary(os.path.join(GOLDEN, "tiny_llama_int8"))      # a decoder-only model is not a Translator model

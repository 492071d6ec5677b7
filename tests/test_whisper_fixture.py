"""Whisper path on the CPU side: the oracle's WhisperOracle against the COMMITTED outputs of the unmodified
reference's models::Whisper (tests/golden/whisper_ref.json, written by tools/make_golden.py --whisper-only) on the tiny
WhisperSpec model of tests/golden/tiny_whisper; plus the host-only model parser of the engine.  No GPU, no /root/reference."""
import json
import os

import numpy as np
import pytest

from oracle import ct2_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL = os.path.join(GOLDEN, "tiny_whisper")


def inputs(seed, batch, n_mels=16, frames=60):
    return (np.random.default_rng(seed).standard_normal((batch, n_mels, frames)) * 2).astype(np.float32)


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "whisper_ref.json")) as f:
        return json.load(f)


def test_conv1d_matches_a_direct_loop():
    r = np.random.default_rng(0)
    x, w, b = r.standard_normal((2, 3, 11)).astype(np.float32), r.standard_normal((4, 3, 3)).astype(np.float32), r.standard_normal(4).astype(np.float32)
    for stride in (1, 2):
        y = O.conv1d(x, w, b, stride, 1)
        tout = (11 + 2 - 3) // stride + 1
        assert y.shape == (2, 4, tout)
        ref = np.zeros_like(y)
        xp = np.pad(x, ((0, 0), (0, 0), (1, 1)))
        for t in range(tout):
            ref[:, :, t] = np.einsum("bck,ock->bo", xp[:, :, t * stride:t * stride + 3], w) + b
        np.testing.assert_allclose(y, ref, atol=1e-5)


def test_oracle_float32_matches_reference_fixture(fixture):
    oracle = O.WhisperOracle.from_dir(MODEL, compute_type="float32")
    ref = fixture["models"]["float32"]
    enc = oracle.encode_features(inputs(ref["encode_seed"], 2))
    np.testing.assert_allclose(enc, np.array(ref["encoder_output"], np.float32), atol=3e-5)
    hyps = 0
    for c in ref["cases"]:
        res, nsp = oracle.generate(inputs(c["seed"], c["batch"]), np.array(c["prompts"]), beam_size=c["beam_size"],
                                   num_hypotheses=c["num_hypotheses"], length_penalty=c["length_penalty"],
                                   max_length=c["max_length"], suppress_blank=c["suppress_blank"])
        np.testing.assert_allclose(nsp, c["no_speech_prob"], rtol=1e-3, atol=1e-7)
        for got, toks, scores in zip(res, c["sequences"], c["scores"]):
            assert [h[0] for h in got] == toks
            np.testing.assert_allclose([h[1] for h in got], scores, atol=1e-4)
            hyps += len(got)
    assert hyps >= 50


def test_translator_summary_parses_the_whisper_model():
    from ctranslate2_b200.translator import translator_summary
    s = translator_summary(MODEL)
    assert s["spec"] == "WhisperSpec" and s["encoder_layers"] == 2 and s["decoder_layers"] == 2 and s["num_heads"] == 4
    assert s["d_model"] == 64 and s["source_vocab"] == 0 and s["target_vocab"] == 122 and s["activation"] == 3
    assert s["embeddings_scale"] == 0 and s["pre_norm"] is True

"""Whisper path on the GPU: ctranslate2_b200.Whisper — Conv1D front-end (im2col + float Dense), GELU encoder,
cross-attention decoder, prompt forwarding and the device-resident search with SuppressTokens / SuppressTokensBegin and the
timestamp rules (ApplyTimestampRules, whisper.cc:742-860) —
through the C-ABI, against the committed outputs of the UNMODIFIED reference's models::Whisper
(tests/golden/whisper_ref.json) and the oracle run live.  float32: every token equal, scores to 2e-4; int8 / float16: the
majority criterion of tests/test_gpu_translator.py (a d = 64 model amplifies single rounding flips)."""
import json
import os

import numpy as np
import pytest

from ctranslate2_b200.whisper import Whisper
from oracle import ct2_oracle as O
from gpu_util import gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL = os.path.join(GOLDEN, "tiny_whisper")


def inputs(seed, batch, n_mels=16, frames=60):
    return (np.random.default_rng(seed).standard_normal((batch, n_mels, frames)) * 2).astype(np.float32)


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "whisper_ref.json")) as f:
        return json.load(f)


def _run(w, c):
    return w.generate(inputs(c["seed"], c["batch"]), c["prompts"], beam_size=c["beam_size"], num_hypotheses=c["num_hypotheses"],
                      length_penalty=c["length_penalty"], max_length=c["max_length"], suppress_blank=c["suppress_blank"],
                      return_scores=True, return_no_speech_prob=True)


@gpu
def test_encoder_matches_reference_and_oracle(fixture):
    w = Whisper(MODEL, compute_type="float32")
    ref = fixture["models"]["float32"]
    x = inputs(ref["encode_seed"], 2)
    enc = w.encode(x)
    np.testing.assert_allclose(enc, np.array(ref["encoder_output"], np.float32), atol=2e-4)
    np.testing.assert_allclose(enc, O.WhisperOracle.from_dir(MODEL, compute_type="float32").encode_features(x), atol=2e-4)
    w.close()


@gpu
def test_float32_generation_equals_the_reference(fixture):
    w = Whisper(MODEL, compute_type="float32")
    total = 0
    for c in fixture["models"]["float32"]["cases"]:
        res = _run(w, c)
        for b, r in enumerate(res):
            assert r.sequences_ids == c["sequences"][b], (c["seed"], c["beam_size"], b)
            np.testing.assert_allclose(r.scores, c["scores"][b], atol=2e-4)
            np.testing.assert_allclose(r.no_speech_prob, c["no_speech_prob"][b], rtol=2e-3, atol=1e-7)
            total += len(r.sequences_ids)
    assert total >= 50
    w.close()


@gpu
@pytest.mark.parametrize("compute", ["int8", "int8_float16", "float16"])
def test_quantized_and_half_generation_agree_statistically(fixture, compute):
    w = Whisper(MODEL, compute_type=compute)
    ref = fixture["models"]["int8" if compute.startswith("int8") else "float32"]
    same = total = 0
    for c in ref["cases"]:
        res = _run(w, c)
        for b, r in enumerate(res):
            total += 1
            same += r.sequences_ids[:1] == c["sequences"][b][:1]
    assert same / total >= 0.6, (same, total)
    w.close()


@gpu
def test_tokens_interface_and_errors():
    w = Whisper(MODEL, compute_type="float32")
    x = inputs(1, 1)
    # the tiny model stores 64 decoder positions: like the reference, positions past the table are an error
    a = w.generate(x, [["<|startoftranscript|>", "<|l0|>", "<|transcribe|>", "<|notimestamps|>"]], beam_size=2, max_length=40)
    b = w.generate(x, [[w.sot_id, w.sot_id + 1, w._ids["<|transcribe|>"], w.no_timestamps_id]], beam_size=2, max_length=40)
    with pytest.raises(ValueError):
        w.generate(x, [[w.sot_id, w.no_timestamps_id]], beam_size=2)          # default max_length 448 > 64 positions
    assert a[0].sequences_ids == b[0].sequences_ids and a[0].sequences[0] == [w._tokens[i] for i in a[0].sequences_ids[0]]
    with pytest.raises(ValueError):
        w.generate(x, [[w.sot_id, w._ids["<|transcribe|>"], 5]], max_length=40)   # text after the task tokens (decoding prefix)
    ts = w.generate(x, [[w.sot_id, w.sot_id + 1, w._ids["<|transcribe|>"]]], beam_size=2, max_length=40)[0].sequences_ids[0]
    assert ts[0] > w.no_timestamps_id                                          # with timestamps the output starts with one
    with pytest.raises(ValueError):
        w.generate(inputs(1, 1, n_mels=8), [[w.sot_id, w.no_timestamps_id]], max_length=40)   # wrong number of mel bins
    with pytest.raises(ValueError):
        w.generate(x, [[w._ids["<|transcribe|>"], w.no_timestamps_id]], max_length=40)        # no <|startoftranscript|>
    assert w.generate(x[:0], []) == []
    w.close()

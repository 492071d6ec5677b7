"""Host logic of the model loader on CPU (no GPU, no compute): ct2b200_model_summary parses model.bin / config.json exactly
as ct2b200_generator_open does (same C++ function) — binary format v6 of the reference (src/models/model.cc:561-660,
python/ctranslate2/specs/model_spec.py:382-414), attribute lookup with defaults, and the geometry of every storage type."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctranslate2_b200 as ct2  # noqa: E402
from ctranslate2_b200.converters.synthetic import LlamaConfig, write_llama_model  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_summary_of_the_model_written_by_the_reference_spec_writer():
    """tests/golden/tiny_llama_int8 was produced by the REFERENCE's own TransformerDecoderModelSpec.save (tools/make_golden.py)."""
    s = ct2.model_summary(os.path.join(GOLDEN, "tiny_llama_int8"))
    assert s["spec"] == "TransformerDecoderSpec" and s["binary_version"] == 6 and s["weights"] == "int8"
    assert s["num_layers"] >= 1 and s["num_heads"] % s["num_heads_kv"] == 0
    assert s["d_model"] == s["num_heads"] * s["head_dim"] and s["rotary_interleave"] is False


@pytest.mark.parametrize("quant,weights", [("int8_float16", "int8"), ("int8", "int8"), ("float16", "float16"),
                                            ("bfloat16", "bfloat16"), ("awq_gemm", "awq_gemm"), ("awq_gemv", "awq_gemv")])
def test_geometry_of_every_storage_type(tmp_path, quant, weights):
    """ffn_dim != d_model and num_heads_kv != num_heads on purpose: an AWQ_GEMM-packed weight is [K, N/8] (the output width is
    NOT its first dimension), an AWQ_GEMV-packed one [N, K/8]."""
    cfg = LlamaConfig(num_layers=3, num_heads=8, num_heads_kv=2, head_dim=64, ffn_dim=1280, vocab_size=320,
                      rotary_scaling_type=2, rotary_scaling_factor=8.0, original_max_position_embeddings=64)
    d = str(tmp_path / quant)
    write_llama_model(d, cfg, quant, seed=3, init_std=0.05)
    s = ct2.model_summary(d)
    assert s["weights"] == weights
    assert (s["num_layers"], s["num_heads"], s["num_heads_kv"], s["head_dim"]) == (3, 8, 2, 64)
    assert (s["d_model"], s["ffn_dim"], s["vocab_size"]) == (512, 1280, 320)
    assert s["rotary_scaling_type"] == 2 and s["rotary_interleave"] is False and s["activation"] == 2
    assert abs(s["rotary_base"] - cfg.rotary_base) < 1e-3 and abs(s["layer_norm_epsilon"] - cfg.rms_eps) < 1e-12


def test_errors(tmp_path):
    with pytest.raises(RuntimeError):
        ct2.model_summary(str(tmp_path / "missing"))
    # a non-decoder spec is rejected with the reference's exception class for bad arguments (std::invalid_argument)
    bad = tmp_path / "bad"
    bad.mkdir()
    import struct
    with open(bad / "model.bin", "wb") as f:
        f.write(struct.pack("I", 6))
        name = b"TransformerSpec\x00"
        f.write(struct.pack("H", len(name)) + name)
        f.write(struct.pack("I", 1))      # revision
        f.write(struct.pack("I", 0))      # no variables
        f.write(struct.pack("I", 0))      # no aliases
    with pytest.raises(ValueError):
        ct2.model_summary(str(bad))


def test_token_ids_are_range_checked_before_they_reach_the_device():
    from ctranslate2_b200.generator import _validate_ids
    _validate_ids([[0, 5, 199], []], 200)
    with pytest.raises(ValueError, match="row 1"):
        _validate_ids([[1, 2], [3, 200]], 200)
    with pytest.raises(ValueError, match="-1"):
        _validate_ids([[1, -1]], 200)


def test_empty_batch_is_an_empty_result_without_touching_the_device():
    from ctranslate2_b200.generator import Generator
    g = Generator.__new__(Generator)        # no model, no device: the empty batch must return before either is needed
    g._h = None
    assert g.generate_batch([]) == []


@pytest.mark.parametrize("extra,omit", [
    ({"decoder/scale_embeddings": __import__("numpy").int8(1)}, ("decoder/scale_embeddings",)),      # sqrt(d_model) scaling on
    ({}, ("decoder/scale_embeddings",)),                                                             # absent = on in the reference
    ({"decoder/scale_embeddings": __import__("numpy").float32(2.5)}, ("decoder/scale_embeddings",)),
    ({"decoder/alibi": __import__("numpy").int8(1)}, ("decoder/alibi",)),
    ({"decoder/sliding_window": __import__("numpy").int32(128)}, ()),
    ({"decoder/layer_0/self_attention/sliding_window": __import__("numpy").int32(128)}, ()),
    ({"decoder/layer_1/self_attention/q_norm/gamma": __import__("numpy").ones(64, "float32")}, ()),
    ({"decoder/layer_1/self_attention/k_norm/gamma": __import__("numpy").ones(64, "float32")}, ()),
    ({"decoder/layer_0/self_attention/layer_norm/layer_norm_use_residual": __import__("numpy").int8(1)}, ()),
    ({"decoder/layer_norm/layer_norm_use_residual": __import__("numpy").int8(1)}, ()),
    ({"decoder/layernorm_embedding/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/project_in/weight": __import__("numpy").ones((512, 512), "float32")}, ()),
    ({"decoder/project_out/weight": __import__("numpy").ones((512, 512), "float32")}, ()),
    ({"decoder/layer_0/input_layer_norm/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/layer_0/post_attention_layer_norm/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/layer_1/pre_feedforward_layer_norm/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/layer_1/post_feedforward_layer_norm/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/layer_2/shared_layer_norm/gamma": __import__("numpy").ones(512, "float32")}, ()),
    ({"decoder/final_logit_softcapping": __import__("numpy").float32(30.0)}, ()),
    ({"decoder/layer_0/self_attention/queries_scale": __import__("numpy").float32(0.5)}, ()),
])
def test_decoder_features_the_engine_does_not_implement_are_refused(tmp_path, extra, omit):
    """The reference honours these TransformerDecoderSpec attributes (transformer.cc:380-400, 475-530, attention.cc:314-315,
    common.cc:448); a model that uses one must not load and silently produce other tokens (Gemma: scaled embeddings + 1 + gamma
    norms, Mistral: sliding window, Qwen3: q/k norms, Falcon / GPT-J: shared / parallel norms)."""
    cfg = LlamaConfig(num_layers=3, num_heads=8, num_heads_kv=2, head_dim=64, ffn_dim=1280, vocab_size=320)
    d = str(tmp_path / "m")
    write_llama_model(d, cfg, "float16", seed=3, init_std=0.05, extra=extra, omit=omit)
    with pytest.raises(ValueError):
        ct2.model_summary(d)


def test_generation_options_are_rejected_unless_neutral():
    """GenerationOptions the engine does not implement raise unless they hold the reference's default — flags are compared
    as flags (True == 1 in Python: disable_unk=True must not pass as "1")."""
    from ctranslate2_b200.generator import _check_options
    _check_options({}, 8, 0)
    _check_options({"repetition_penalty": 1.0, "no_repeat_ngram_size": 0, "disable_unk": False, "suppress_sequences": [],
                    "static_prompt": None, "sampling_temperature": 1, "num_hypotheses": 1, "asynchronous": False,
                    "cache_static_prompt": True, "callback": None}, 8, 8)
    for bad in ({"disable_unk": True}, {"return_logits_vocab": True}, {"return_alternatives": True}, {"asynchronous": True},
                {"no_repeat_ngram_size": 1}, {"repetition_penalty": 1.2}, {"sampling_temperature": 0.7}, {"num_hypotheses": 2},
                {"suppress_sequences": [["a"]]}, {"static_prompt": ["a"]}, {"sampling_topp": 0.9}, {"patience": 2},
                {"cache_static_prompt": False}, {"not_an_option": 0}, {"num_hypotheses": True}):
        with pytest.raises(ValueError):
            _check_options(bad, 8, 0)
    with pytest.raises(ValueError):
        _check_options({}, 0, 0)            # max_length == 0 (decoding.cc:1035-1040)
    with pytest.raises(ValueError):
        _check_options({}, 4, 5)            # min_length > max_length

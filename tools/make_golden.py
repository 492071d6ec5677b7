#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the reference itself.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/Makefile.ref):

    python tools/make_golden.py

1. tests/golden/ref_gtest_vectors.json — the golden vectors the reference's own gtests hold for
   this path, extracted verbatim (by parsing, not by hand) from /root/reference/tests/ops_test.cc and
   layers_test.cc: every `StorageView name({shape}, std::vector<T>{...})` inside the named TEST_P blocks.
2. tests/golden/tiny_llama_int8/ — a 2-layer GQA/SwiGLU/RoPE decoder written by the REFERENCE's
   python spec writer (python/ctranslate2/specs), quantization="int8".
3. tests/golden/tiny_llama_int8_ref.npz — outputs of the UNMODIFIED reference (oracle/_ref, CPU):
   forward logits for a seeded prompt batch and greedy generate_batch tokens.
4. tests/golden/ref_ops_random.npz — reference op outputs (Quantize, Gemm s8, Dequantize, RMSNorm,
   Rotary, SoftMax, TopK, Gather) on seeded random inputs.

Nothing here runs at test time on the GPU box; the tests read only the files this script wrote.
"""
import json
import os
import re
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

GTESTS = {
    "ops_test.cc": ["GemmInt8", "TopK", "TopKVariableDepth", "TopKChangeK", "SoftMax", "LogSoftMax",
                    "MaskedSoftMax", "RMSNorm", "LayerNorm", "QuantizeINT8", "QuantizeINT8ZeroRow", "Swish", "ReLU",
                    "GELU", "GELUTanh", "GELUSigmoid", "Gemm", "GemmBias", "GemmResidual", "GemmGELU",
                    "GatherData1D", "GatherData1DIndex2D", "GatherData2D", "GatherData3D",
                    "GatherData2DIndex2D", "BiasAddGELU", "BiasAddAxisGELU"],
    "layers_test.cc": ["RotaryEmbedding"],
}


STATICS = ("gemm_a", "gemm_b", "gemm_y", "bias_value", "bias_bias")


def extract_gtest_vectors():
    out = {}
    sv = re.compile(r"StorageView\s+(\w+)\s*(?:=\s*StorageView)?\((\{[^}]*\}|\w+\.shape\(\)),\s*std::vector<(\w+)>\s*\{([^}]*)\}", re.S)
    for fname, names in GTESTS.items():
        text = open(os.path.join(REF, "tests", fname)).read()
        for name in names:
            m = re.search(r"TEST_P\(\w+,\s*%s\)\s*\{" % re.escape(name), text)
            if not m:
                print("  (no such test in this reference version: %s)" % name)
                continue
            nxt = re.search(r"\nTEST(_P|_F)?\(", text[m.end():])
            body = text[m.end(): m.end() + (nxt.start() if nxt else len(text))]
            items = []
            for v in sv.finditer(body):
                if v.group(2).startswith("{"):
                    shape = [int(s) for s in v.group(2).strip("{}").split(",") if s.strip()]
                else:   # `other.shape()`: same shape as the named vector of this test
                    other = v.group(2).split(".")[0]
                    shape = next(i["shape"] for i in items if i["name"] == other)
                vals = [float(t.rstrip("f")) for t in re.split(r"[,\s]+", v.group(4).strip()) if t]
                items.append({"name": v.group(1), "shape": shape, "ctype": v.group(3), "values": vals})
            out[name] = items
            print("  %-24s %d vectors" % (name, len(items)))
        # file-scope inputs shared by several tests (`static const StorageView gemm_a(...)`)
        static = re.compile(r"static const StorageView\s+(\w+)\((\{[^}]*\}),\s*std::vector<(\w+)>\s*\{([^}]*)\}", re.S)
        for v in static.finditer(text):
            if v.group(1) in STATICS:
                shape = [int(t) for t in v.group(2).strip("{}").split(",") if t.strip()]
                vals = [float(t.rstrip("f")) for t in re.split(r"[,\s]+", v.group(4).strip()) if t]
                out.setdefault("_static", []).append({"name": v.group(1), "shape": shape, "ctype": v.group(3), "values": vals})
                print("  static %-17s %s" % (v.group(1), shape))
    return out


def make_tiny_model(model_dir):
    sys.path.insert(0, os.path.join(REF, "python"))
    from ctranslate2.specs import common_spec, transformer_spec
    L, H, Hkv, D, F, V = 2, 4, 2, 32, 256, 200
    d = H * D
    spec = transformer_spec.TransformerDecoderModelSpec.from_config(
        L, H, activation=common_spec.Activation.SWISH, pre_norm=True, ffn_glu=True, rms_norm=True,
        rotary_dim=0, rotary_interleave=False, rotary_base=500000.0, num_heads_kv=Hkv)
    rng = np.random.default_rng(1234)

    def lin(s, n, k):
        s.weight = (rng.standard_normal((n, k)) * 0.05).astype(np.float32)

    dec = spec.decoder
    dec.scale_embeddings = False
    dec.embeddings.weight = (rng.standard_normal((V, d)) * 0.05).astype(np.float32)
    dec.layer_norm.gamma = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    for l in dec.layer:
        l.self_attention.layer_norm.gamma = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        l.ffn.layer_norm.gamma = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        lin(l.self_attention.linear[0], d + 2 * Hkv * D, d)
        lin(l.self_attention.linear[1], d, d)
        lin(l.ffn.linear_0, F, d)
        lin(l.ffn.linear_0_noact, F, d)
        lin(l.ffn.linear_1, d, F)
    lin(dec.projection, V, d)
    spec.register_vocabulary(["<t%d>" % i for i in range(V)])
    spec.config.bos_token = "<t1>"
    spec.config.eos_token = "<t2>"
    spec.config.unk_token = "<t0>"
    spec.config.layer_norm_epsilon = 1e-5
    spec.validate()
    spec.optimize(quantization="int8")
    shutil.rmtree(model_dir, ignore_errors=True)
    os.makedirs(model_dir)
    spec.save(model_dir)
    return V


def make_scores_fixture():
    """GenerationResult.scores of the unmodified reference on the committed tiny model (return_scores=true): sum of the
    chosen tokens' log-probabilities / length^length_penalty, incl. rows that end on the end token."""
    from oracle import refapi
    assert refapi.available(), "build oracle/_ref first: make -f oracle/Makefile.ref -j8"
    mdir = os.path.join(OUT, "tiny_llama_int8")
    fx = np.load(os.path.join(OUT, "tiny_llama_int8_ref.npz"), allow_pickle=True)
    prompts = fx["prompts"]
    g = refapi.RefGenerator(mdir, "int8", 4)
    gm = fx["generated_min12"]
    # end tokens that rows emit at different positions: rows then stop at step 0, mid-sequence, or never
    ends = [2, int(gm[0][3]), int(gm[1][2]), int(gm[2][5])]
    cases = []
    for lp in (1.0, 0.0, 0.6):
        for end in ends:
            for (mx, mn) in ((12, 12), (12, 0), (12, 3)):
                toks, scores = g.generate_with_scores(prompts, mx, mn, end, lp)
                cases.append({"max_length": mx, "min_length": mn, "end_id": end, "length_penalty": lp, "tokens": toks,
                              "scores": [float(x) for x in scores]})
    beams = []
    for beam in (2, 4):
        for lp in (1.0, 0.0):
            for end in ends[:3]:
                for (mx, mn, nh, pat) in ((10, 0, 2, 1.0), (10, 3, 2, 2.0), (6, 6, 1, 1.0)):
                    r = g.generate_beam(prompts, beam, mx, mn, end, lp, nh, pat)
                    beams.append({"beam_size": beam, "max_length": mx, "min_length": mn, "end_id": end,
                                  "length_penalty": lp, "num_hypotheses": nh, "patience": pat,
                                  "hypotheses": [[[t, sc] for t, sc in row] for row in r]})
    g.close()
    with open(os.path.join(OUT, "tiny_llama_int8_scores.json"), "w") as f:
        json.dump({"prompts": prompts.tolist(), "cases": cases, "beam_cases": beams}, f)
    print("wrote tiny_llama_int8_scores.json (%d greedy cases, %d beam cases)" % (len(cases), len(beams)))


def make_processors_fixture():
    """Greedy generate_batch of the unmodified reference with the logits processors of GenerationOptions (repetition_penalty,
    no_repeat_ngram_size, disable_unk, suppress_sequences) on the committed tiny model — it repeats tokens a lot, which is
    exactly what these options act on."""
    from oracle import refapi
    assert refapi.available(), "build oracle/_ref first: make -f oracle/Makefile.ref -j8"
    mdir = os.path.join(OUT, "tiny_llama_int8")
    fx = np.load(os.path.join(OUT, "tiny_llama_int8_ref.npz"), allow_pickle=True)
    prompts = fx["prompts"]
    gm = fx["generated_min12"]
    g = refapi.RefGenerator(mdir, "int8", 4)
    options = [dict(repetition_penalty=1.3), dict(repetition_penalty=0.7), dict(repetition_penalty=2.0, no_repeat_ngram_size=3),
               dict(no_repeat_ngram_size=1), dict(no_repeat_ngram_size=2), dict(no_repeat_ngram_size=4), dict(disable_unk=True),
               dict(suppress_sequences=[[int(gm[0][0])]]),
               dict(suppress_sequences=[[int(gm[0][0])], [int(gm[1][0]), int(gm[1][1])], [int(gm[2][1]), int(gm[2][2]), int(gm[2][3])]]),
               dict(suppress_sequences=[[int(gm[1][4]), int(gm[1][5])]], repetition_penalty=1.2, no_repeat_ngram_size=2)]
    cases = []
    for opt in options:
        for (mx, mn, end) in ((12, 12, 2), (12, 0, int(gm[0][4])), (10, 3, int(gm[1][2]))):
            toks, scores = g.generate_processors(prompts, mx, mn, end, **opt)
            cases.append({"max_length": mx, "min_length": mn, "end_id": end, "options": opt, "tokens": toks,
                          "scores": [float(x) for x in scores]})
    g.close()
    with open(os.path.join(OUT, "tiny_llama_int8_processors.json"), "w") as f:
        json.dump({"prompts": prompts.tolist(), "cases": cases}, f)
    print("wrote tiny_llama_int8_processors.json (%d cases)" % len(cases))


def make_ragged_fixture():
    """Greedy generate_batch of the unmodified reference over prompts of different lengths (include_prompt_in_result=false):
    the shortest prompt decides how much is forwarded at once, the rest of each prompt is forced through the loop
    (language_model.cc:217-238).  Also records the reference's behaviour when the shortest prompt is ONE token: it then keeps
    return_prefix = true and returns the forced prompt tokens as part of the result."""
    from oracle import refapi
    assert refapi.available(), "build oracle/_ref first: make -f oracle/Makefile.ref -j8"
    g = refapi.RefGenerator(os.path.join(OUT, "tiny_llama_int8"), "int8", 4)
    batches = [[[5, 9, 11, 40, 7], [8, 3, 77], [100, 23, 45, 67]],
               [[17, 4], [9, 9, 9, 9, 9, 9, 9], [150, 3, 8]],
               [[5, 9, 11, 40, 7], [8], [100, 23]],
               [[5], [8], [100]]]
    cases = []
    for prompts in batches:
        for (mx, mn, end) in ((6, 6, 2), (6, 0, 164), (5, 2, 18), (8, 3, 143)):
            toks, scores = g.generate_ragged(prompts, mx, mn, end)
            cases.append({"prompts": prompts, "max_length": mx, "min_length": mn, "end_id": end, "tokens": toks,
                          "scores": [float(x) for x in scores]})
    g.close()
    with open(os.path.join(OUT, "tiny_llama_int8_ragged.json"), "w") as f:
        json.dump({"cases": cases}, f)
    print("wrote tiny_llama_int8_ragged.json (%d cases)" % len(cases))


def make_score_fixture():
    """Generator::score_batch of the unmodified reference on the tiny model: log-probability of every token given its prefix,
    ragged batch, sequences too short to score, ScoringOptions::offset."""
    from oracle import refapi
    assert refapi.available(), "build oracle/_ref first: make -f oracle/Makefile.ref -j8"
    g = refapi.RefGenerator(os.path.join(OUT, "tiny_llama_int8"), "int8", 4)
    rng = np.random.default_rng(11)
    seqs = [[int(t) for t in rng.integers(3, 200, size=n)] for n in (12, 2, 1, 30, 7, 19)]
    cases = [{"sequences": seqs, "offset": off, "log_probs": g.score(seqs, off)} for off in (0, 1, 5)]
    g.close()
    with open(os.path.join(OUT, "tiny_llama_int8_score_batch.json"), "w") as f:
        json.dump({"cases": cases}, f)
    print("wrote tiny_llama_int8_score_batch.json (%d cases)" % len(cases))


SEQ2SEQ_CASES = [  # (beam, num_hypotheses, length_penalty, max_length, min_length)
    (1, 1, 1.0, 16, 1), (2, 2, 1.0, 16, 1), (4, 2, 0.0, 16, 1), (3, 3, 0.6, 12, 1), (4, 4, 1.0, 14, 9), (2, 1, 1.0, 5, 1)]


def seq2seq_sources(seed, cases, lo, hi):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(cases):
        batch = int(rng.integers(1, 5))
        out.append([[int(x) for x in rng.integers(lo, hi, size=int(rng.integers(2, 10)))] for _ in range(batch)])
    return out


def make_seq2seq_fixture():
    """Encoder-decoder path: outputs of the UNMODIFIED reference's Translator (oracle/_ref, CPU) on
    (a) the reference's own golden model tests/data/models/v2/aren-transliteration-i8 (tests/translator_test.cc:53-96), copied
        to tests/golden/ as a data fixture (binary version 2: the activation quantizer truncates), and
    (b) a post-norm / Swish / start-from-zero-embedding model written by converters/synthetic.py (the OPUS-MT recipe in small),
    in float32 (no activation quantization: every token must agree) and int8."""
    from ctranslate2_b200.converters.synthetic import TransformerConfig, write_transformer_model
    from oracle import refapi
    src_dir = os.path.join(REF, "tests", "data", "models", "v2", "aren-transliteration-i8")
    aren = os.path.join(OUT, "aren-transliteration-i8")
    if os.path.isdir(aren):
        shutil.rmtree(aren)
    shutil.copytree(src_dir, aren)
    for root, _, files in os.walk(aren):          # the reference tree is read-only: the copy must not be
        os.chmod(root, 0o755)
        for name in files:
            os.chmod(os.path.join(root, name), 0o644)
    post = os.path.join(OUT, "tiny_seq2seq_postnorm")
    cfg = TransformerConfig(encoder_layers=2, decoder_layers=2, num_heads=4, d_model=64, ffn_dim=128, source_vocab=120,
                            target_vocab=96, pre_norm=False, activation=2, start_from_zero_embedding=True)
    write_transformer_model(post, cfg, "int8", seed=7)
    fixture = {}
    for name, mdir, lo, hi in (("aren", aren, 4, 51), ("postnorm", post, 3, 120)):
        entry = {"models": {}}
        for compute in ("float32", "int8"):
            t = refapi.RefTranslator(mdir, compute, 2)
            cases = []
            for ci, (beam, nh, lp, mx, mn) in enumerate(SEQ2SEQ_CASES):
                for srcs in seq2seq_sources(100 + ci, 4, lo, hi):
                    res = t.translate(srcs, beam_size=beam, num_hypotheses=nh, max_length=mx, min_length=mn, length_penalty=lp)
                    cases.append({"sources": srcs, "beam_size": beam, "num_hypotheses": nh, "length_penalty": lp,
                                  "max_length": mx, "min_length": mn,
                                  "hypotheses": [[h[0] for h in r] for r in res], "scores": [[h[1] for h in r] for r in res]})
            srcs = seq2seq_sources(9, 1, lo, hi)[0] + [[lo + 1, lo + 2, lo + 3, lo + 4, lo + 5, lo + 6, lo + 7]]
            memory, lens = t.encode(srcs)
            d = 32 if name == "aren" else 64
            S = max(len(r) for r in srcs)
            memory = memory.reshape(-1)[:len(srcs) * S * d].reshape(len(srcs), S, d)
            entry["models"][compute] = {"cases": cases, "encode_sources": srcs, "memory": memory.tolist()}
            t.close()
        fixture[name] = entry
    with open(os.path.join(OUT, "seq2seq_ref.json"), "w") as f:
        json.dump(fixture, f)
    print("seq2seq fixture:", sum(len(m["cases"]) for e in fixture.values() for m in e["models"].values()), "cases")


WHISPER_CASES = [  # (beam, num_hypotheses, length_penalty, max_length, suppress_blank, timestamps)
    (1, 1, 1.0, 24, True, False), (3, 2, 1.0, 24, True, False), (5, 3, 1.0, 30, True, False), (5, 1, 0.0, 24, False, False),
    (2, 2, 0.7, 16, True, False), (1, 1, 1.0, 30, True, True), (5, 2, 1.0, 30, True, True), (3, 3, 1.0, 24, False, True)]


def whisper_inputs(seed, batch, n_mels, frames):
    return (np.random.default_rng(seed).standard_normal((batch, n_mels, frames)) * 2).astype(np.float32)


def make_whisper_fixture():
    """Whisper path: outputs of the UNMODIFIED reference's models::Whisper (oracle/_ref, CPU) on a tiny WhisperSpec
    model written by converters/synthetic.py (2 + 2 layers, d 64, 4 heads, 16 mel bins, 60 frames): encoder output, generate
    (greedy and beam, several hypotheses, length penalties, suppress_blank on / off) and no-speech probabilities, in float32
    and int8.  The features are re-generated from their seeds by the tests."""
    from ctranslate2_b200.converters.synthetic import WhisperConfig, whisper_vocabulary, write_whisper_model
    from oracle import refapi
    cfg = WhisperConfig(encoder_layers=2, decoder_layers=2, num_heads=4, d_model=64, n_mels=16, max_source_positions=30,
                        max_target_positions=64, text_tokens=100, languages=3, timestamps=11)
    mdir = os.path.join(OUT, "tiny_whisper")
    write_whisper_model(mdir, cfg, "int8", seed=5)
    vocab = whisper_vocabulary(cfg)
    sot = vocab.index("<|startoftranscript|>")
    fixture = {"n_mels": 16, "frames": 60, "d_model": 64, "models": {}}
    for compute in ("float32", "int8"):
        w = refapi.RefWhisper(mdir, compute, 2)
        cases = []
        for ci, (beam, nh, lp, mx, blank, stamps) in enumerate(WHISPER_CASES):
            for rep in range(3):
                seed, batch = 300 + 10 * ci + rep, 1 + (ci + rep) % 4
                feats = whisper_inputs(seed, batch, 16, 60)
                # with timestamps the prompt ends with the task token and ApplyTimestampRules shapes the output
                prompts = [[sot, sot + 1 + (b % 3), vocab.index("<|transcribe|>" if b % 2 == 0 else "<|translate|>")] +
                           ([] if stamps else [vocab.index("<|notimestamps|>")]) for b in range(batch)]
                res, nsp = w.generate(feats, prompts, beam_size=beam, num_hypotheses=nh, length_penalty=lp, max_length=mx,
                                      suppress_blank=blank)
                cases.append({"seed": seed, "batch": batch, "prompts": prompts, "beam_size": beam, "num_hypotheses": nh,
                              "length_penalty": lp, "max_length": mx, "suppress_blank": blank, "timestamps": stamps,
                              "sequences": [[h[0] for h in r] for r in res], "scores": [[h[1] for h in r] for r in res],
                              "no_speech_prob": [float(x) for x in nsp]})
        enc = w.encode(whisper_inputs(7, 2, 16, 60), 64)
        fixture["models"][compute] = {"cases": cases, "encode_seed": 7, "encoder_output": enc.tolist()}
        w.close()
    with open(os.path.join(OUT, "whisper_ref.json"), "w") as f:
        json.dump(fixture, f)
    print("whisper fixture:", sum(len(m["cases"]) for m in fixture["models"].values()), "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--scores-only" in sys.argv:
        make_scores_fixture()
        return
    if "--score-only" in sys.argv:
        make_score_fixture()
        return
    if "--ragged-only" in sys.argv:
        make_ragged_fixture()
        return
    if "--whisper-only" in sys.argv:
        make_whisper_fixture()
        return
    if "--seq2seq-only" in sys.argv:
        make_seq2seq_fixture()
        return
    if "--processors-only" in sys.argv:
        make_processors_fixture()
        return
    if "--gtest-only" in sys.argv:
        with open(os.path.join(OUT, "ref_gtest_vectors.json"), "w") as f:
            json.dump(extract_gtest_vectors(), f)
        return
    print("extracting gtest golden vectors")
    with open(os.path.join(OUT, "ref_gtest_vectors.json"), "w") as f:
        json.dump(extract_gtest_vectors(), f)

    from oracle import refapi
    assert refapi.available(), "build oracle/_ref first: make -f oracle/Makefile.ref -j8"

    print("tiny model via the reference spec writer")
    mdir = os.path.join(OUT, "tiny_llama_int8")
    V = make_tiny_model(mdir)
    g = refapi.RefGenerator(mdir, "int8", 4)
    rng = np.random.default_rng(42)
    prompts = rng.integers(3, V, size=(3, 7), dtype=np.int32)
    logits = g.forward(prompts)
    gen = g.generate(prompts, max_length=12, min_length=0, end_id=2)
    gen_min = g.generate(prompts, max_length=12, min_length=12, end_id=2)
    np.savez(os.path.join(OUT, "tiny_llama_int8_ref.npz"), prompts=prompts, logits=logits,
             generated=np.array(gen, dtype=object), generated_min12=np.array(gen_min, dtype=np.int32),
             allow_pickle=True)
    g.close()

    print("reference ops on seeded random inputs")
    r = np.random.default_rng(7)
    d = {}
    x = (r.standard_normal((5, 96)) * 3).astype(np.float32)
    x[3] = 0
    d["q_x"] = x
    d["q_q"], d["q_s"] = refapi.quantize(x)
    a = r.integers(-127, 128, size=(5, 96), dtype=np.int8)
    b = r.integers(-127, 128, size=(24, 96), dtype=np.int8)
    d["g_a"], d["g_b"], d["g_c"] = a, b, refapi.gemm_s8(a, b)
    sa = r.uniform(1, 50, 5).astype(np.float32)
    sb = r.uniform(100, 900, 24).astype(np.float32)
    bias = r.standard_normal(24).astype(np.float32)
    d["dq_sa"], d["dq_sb"], d["dq_bias"] = sa, sb, bias
    for act in (-1, 0, 1, 2, 3, 4, 5, 6):
        d["dq_y_act%d" % act] = refapi.dequantize_gemm(d["g_c"], sa, sb, bias, act)
    d["dq_y_nobias"] = refapi.dequantize_gemm(d["g_c"], sa, sb, None, -1)
    gamma = r.standard_normal(96).astype(np.float32)
    d["rn_gamma"], d["rn_y"] = gamma, refapi.rms_norm(gamma, x, 1e-5)
    xr = r.standard_normal((2, 3, 4, 16)).astype(np.float32)
    ang = r.standard_normal((4, 16)).astype(np.float32)
    d["ro_x"], d["ro_sin"], d["ro_cos"] = xr, np.sin(ang), np.cos(ang)
    d["ro_y_interleave"] = refapi.rotary(xr, d["ro_sin"], d["ro_cos"], True)
    d["ro_y_half"] = refapi.rotary(xr, d["ro_sin"], d["ro_cos"], False)
    sx = r.standard_normal((6, 33)).astype(np.float32)
    lens = np.array([33, 1, 7, 20, 33, 2], np.int32)
    d["sm_x"], d["sm_len"] = sx, lens
    d["sm_y"], d["sm_y_len"] = refapi.softmax(sx), refapi.softmax(sx, lens)
    d["sm_logy"] = refapi.softmax(sx, None, True)
    tx = r.standard_normal((4, 1000)).astype(np.float32)
    tx[1, 17] = tx[1, 500] = 9.0   # exact tie: lowest index must win
    d["tk_x"] = tx
    d["tk_v1"], d["tk_i1"] = refapi.topk(tx, 1)
    d["tk_v4"], d["tk_i4"] = refapi.topk(tx, 4)
    gd = r.standard_normal((50, 8)).astype(np.float32)
    gi = r.integers(0, 50, 9).astype(np.int32)
    d["ga_d"], d["ga_i"], d["ga_y"] = gd, gi, refapi.gather(gd, gi)
    np.savez(os.path.join(OUT, "ref_ops_random.npz"), **d)
    make_scores_fixture()
    make_processors_fixture()
    make_ragged_fixture()
    make_score_fixture()
    make_seq2seq_fixture()
    make_whisper_fixture()
    print("done")


if __name__ == "__main__":
    main()

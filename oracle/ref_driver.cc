// oracle/ref_driver.cc — TEST INFRASTRUCTURE (checker / CPU baseline), never shipped.
//
// A thin C-ABI over the UNMODIFIED reference library (oracle/_ref/libct2ref.so,
// built by oracle/Makefile.ref from /root/reference).  It lets the Python tests,
// tools/make_golden.py and bench.py's `--impl reference` / `cpu_baseline` legs call
// the reference's own public C++ API:
//   ctranslate2::Generator::generate_batch_async   include/ctranslate2/generator.h:14-18
//   ctranslate2::Generator::forward_batch_async    include/ctranslate2/generator.h:30-32
//   ctranslate2::ops::{Quantize,Gemm,Dequantize,RMSNorm,Rotary,SoftMax,TopK,Gather}
// Only `tests/`, `__graft_entry__.smoke()` and bench.py's CPU legs may load this.
//
// Built twice: against oracle/_ref (CPU-only reference) and, with -DREF_DRIVER_CUDA, against
// oracle/_ref_cuda (the reference WITH its CUDA backend, oracle/Makefile.ref_cuda).  In the second
// build `ref_set_device(1, flash)` makes the generator / translator entry points load the model on
// Device::CUDA, and the `ref_cuda_*` entry points run the reference's GPU-only ops (AWQ) and its
// CUDA specialisations of the row ops on device buffers staged from host arrays.

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <ctranslate2/generator.h>
#include <ctranslate2/models/language_model.h>
#include <ctranslate2/models/sequence_to_sequence.h>
#include <ctranslate2/models/whisper.h>
#include <ctranslate2/translator.h>
#include <ctranslate2/ops/ops.h>
#include <ctranslate2/utils.h>

using namespace ctranslate2;

namespace {
  thread_local std::string g_error;
  Device g_device = Device::CPU;
  bool g_flash_attention = false;

  struct RefGenerator {
    std::shared_ptr<const models::Model> model;
    std::unique_ptr<Generator> generator;
    const Vocabulary* vocab = nullptr;
  };

  template <typename F>
  int guarded(F&& f) {
    try {
      f();
      return 0;
    } catch (const std::exception& e) {
      g_error = e.what();
      return 1;
    }
  }

  StorageView view_f32(const float* p, Shape shape) {
    return StorageView(std::move(shape), const_cast<float*>(p), Device::CPU);
  }
}

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// 0 = Device::CPU (default), 1 = Device::CUDA (only in the -DREF_DRIVER_CUDA build); flash_attention selects
// layers::FlashMultiHeadAttention (models::ModelLoader::use_flash_attention, transformer.cc:157-169)
int ref_set_device(int cuda, int flash_attention) {
  return guarded([&] {
#ifndef REF_DRIVER_CUDA
    if (cuda) throw std::runtime_error("this driver was built against the CPU-only reference");
#endif
    g_device = cuda ? Device::CUDA : Device::CPU;
    g_flash_attention = flash_attention != 0;
  });
}

void* ref_generator_open(const char* model_dir, const char* compute_type, int intra_threads) {
  RefGenerator* g = nullptr;
  int rc = guarded([&] {
    auto holder = std::make_unique<RefGenerator>();
    models::ModelLoader loader(model_dir);
    loader.device = g_device;
    loader.device_indices = {0};
    loader.compute_type = str_to_compute_type(compute_type);
    loader.use_flash_attention = g_flash_attention;
    holder->model = loader.load().at(0);
    ReplicaPoolConfig config;
    config.num_threads_per_replica = intra_threads > 0 ? intra_threads : 0;
    holder->generator = std::make_unique<Generator>(holder->model, config);
    auto lm = dynamic_cast<const models::LanguageModel*>(holder->model.get());
    if (!lm)
      throw std::runtime_error("model is not a language model");
    holder->vocab = &lm->get_vocabulary();
    g = holder.release();
  });
  return rc == 0 ? g : nullptr;
}

void ref_generator_close(void* handle) { delete static_cast<RefGenerator*>(handle); }

// Greedy generate_batch (beam 1, sampling_topk 1, include_prompt_in_result=false,
// end_token = `end_id` (pass an id that is never the argmax to force max_len tokens)).
// prompt_ids [B,P] int32, out_ids [B,max_len] int32 (padded with -1), out_lens [B].
int ref_generate(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int min_len,
                 int end_id, int32_t* out_ids, int32_t* out_lens) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b) {
      prompts[b].reserve(P);
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    }
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = min_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = false;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    auto futures = g->generator->generate_batch_async(prompts, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      const auto& ids = result.sequences_ids.at(0);
      out_lens[b] = static_cast<int32_t>(ids.size());
      for (int t = 0; t < max_len; ++t)
        out_ids[b * max_len + t] = t < (int)ids.size() ? (int32_t)ids[t] : -1;
    }
  });
}

// Same with GenerationOptions::return_scores = true and a length penalty: out_scores [B] = the reference's
// GenerationResult::scores[0] (sum of the chosen tokens' log-probabilities / length^length_penalty, decoding.cc:189-203).
int ref_generate_scores(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int min_len, int end_id,
                        float length_penalty, int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = min_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = true;
    opt.length_penalty = length_penalty;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    auto futures = g->generator->generate_batch_async(prompts, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      const auto& ids = result.sequences_ids.at(0);
      out_lens[b] = static_cast<int32_t>(ids.size());
      out_scores[b] = result.scores.at(0);
      for (int t = 0; t < max_len; ++t)
        out_ids[b * max_len + t] = t < (int)ids.size() ? (int32_t)ids[t] : -1;
    }
  });
}

// Beam search (GenerationOptions::beam_size > 1, BeamSearch::search, decoding.cc:425-720): the first `num_hyp` hypotheses per
// prompt, best first.  out_ids [B, num_hyp, max_len] (-1 padded), out_lens / out_scores [B, num_hyp].
int ref_generate_beam(void* handle, const int32_t* prompt_ids, int B, int P, int beam_size, int num_hyp, int max_len,
                      int min_len, int end_id, float length_penalty, float patience, int32_t* out_ids, int32_t* out_lens,
                      float* out_scores) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = beam_size;
    opt.patience = patience;
    opt.num_hypotheses = num_hyp;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = min_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = true;
    opt.length_penalty = length_penalty;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    auto futures = g->generator->generate_batch_async(prompts, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      for (int h = 0; h < num_hyp; ++h) {
        const bool have = h < static_cast<int>(result.sequences_ids.size());
        const int64_t o = (static_cast<int64_t>(b) * num_hyp + h);
        out_lens[o] = have ? static_cast<int32_t>(result.sequences_ids[h].size()) : -1;
        out_scores[o] = have ? result.scores.at(h) : 0.f;
        for (int t = 0; t < max_len; ++t)
          out_ids[o * max_len + t] = (have && t < out_lens[o]) ? static_cast<int32_t>(result.sequences_ids[h][t]) : -1;
      }
    }
  });
}

// ---- Translator (encoder-decoder) over token ids: the golden case of tests/translator_test.cc:53-96 --------------------
struct RefTranslator {
  std::shared_ptr<const models::Model> model;
  std::unique_ptr<Translator> translator;
  const Vocabulary* source = nullptr;
  const Vocabulary* target = nullptr;
};

void* ref_translator_open(const char* model_dir, const char* compute_type, int intra_threads) {
  RefTranslator* t = nullptr;
  int rc = guarded([&] {
    auto holder = std::make_unique<RefTranslator>();
    models::ModelLoader loader(model_dir);
    loader.device = g_device;
    loader.device_indices = {0};
    loader.compute_type = str_to_compute_type(compute_type);
    holder->model = loader.load().at(0);
    ReplicaPoolConfig config;
    config.num_threads_per_replica = intra_threads > 0 ? intra_threads : 0;
    holder->translator = std::make_unique<Translator>(holder->model, config);
    auto s2s = dynamic_cast<const models::SequenceToSequenceModel*>(holder->model.get());
    if (!s2s)
      throw std::runtime_error("model is not a sequence-to-sequence model");
    holder->source = &s2s->get_source_vocabulary();
    holder->target = &s2s->get_target_vocabulary();
    t = holder.release();
  });
  return rc == 0 ? t : nullptr;
}

void ref_translator_close(void* handle) { delete static_cast<RefTranslator*>(handle); }

int ref_translator_vocab_sizes(void* handle, int* source, int* target) {
  auto* t = static_cast<RefTranslator*>(handle);
  return guarded([&] {
    *source = static_cast<int>(t->source->size());
    *target = static_cast<int>(t->target->size());
  });
}

// source_ids [B, S] (rows right-padded with -1), beam search with TranslationOptions defaults except the given ones.
// out_ids [B, num_hyp, max_len] (-1 padded), out_lens / out_scores [B, num_hyp].
int ref_translate(void* handle, const int32_t* source_ids, int B, int S, int beam_size, int num_hyp, int max_len,
                  int min_len, float length_penalty, int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  auto* t = static_cast<RefTranslator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> source(B);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < S && source_ids[b * S + i] >= 0; ++i)
        source[b].push_back(t->source->to_token(source_ids[b * S + i]));
    TranslationOptions opt;
    opt.beam_size = beam_size;
    opt.num_hypotheses = num_hyp;
    opt.max_decoding_length = max_len;
    opt.min_decoding_length = min_len;
    opt.length_penalty = length_penalty;
    opt.return_scores = true;
    auto results = t->translator->translate_batch(source, opt);
    for (int b = 0; b < B; ++b) {
      for (int h = 0; h < num_hyp; ++h) {
        const bool have = h < static_cast<int>(results[b].hypotheses.size());
        const int64_t o = static_cast<int64_t>(b) * num_hyp + h;
        out_lens[o] = have ? static_cast<int32_t>(results[b].hypotheses[h].size()) : -1;
        out_scores[o] = have ? results[b].scores.at(h) : 0.f;
        for (int i = 0; i < max_len; ++i)
          out_ids[o * max_len + i] = (have && i < out_lens[o]) ? static_cast<int32_t>(t->target->to_id(results[b].hypotheses[h][i])) : -1;
      }
    }
  });
}

// ---- Whisper (include/ctranslate2/models/whisper.h:86-190) ----
struct RefWhisper {
  std::unique_ptr<models::Whisper> pool;
};

void* ref_whisper_open(const char* model_dir, const char* compute_type, int intra_threads) {
  RefWhisper* w = nullptr;
  int rc = guarded([&] {
    auto holder = std::make_unique<RefWhisper>();
    ReplicaPoolConfig config;
    config.num_threads_per_replica = intra_threads > 0 ? intra_threads : 0;
    holder->pool = std::make_unique<models::Whisper>(model_dir, g_device, str_to_compute_type(compute_type),
                                                     std::vector<int>{0}, /*tensor_parallel=*/false, config);
    w = holder.release();
  });
  return rc == 0 ? w : nullptr;
}
void ref_whisper_close(void* handle) { delete static_cast<RefWhisper*>(handle); }

// features [B, n_mels, T] fp32 -> encoder output [B, T / 2, d] fp32 (out_capacity floats)
int ref_whisper_encode(void* handle, const float* features, int B, int n_mels, int T, float* out, int64_t out_capacity) {
  auto* w = static_cast<RefWhisper*>(handle);
  return guarded([&] {
    StorageView f = view_f32(features, {B, n_mels, T});
    StorageView enc = w->pool->encode(f, /*to_cpu=*/true).get();
    StorageView e32 = enc.to_float32();
    if (e32.size() > out_capacity) throw std::runtime_error("ref_whisper_encode: output buffer too small");
    std::memcpy(out, e32.data<float>(), e32.size() * sizeof(float));
  });
}

// prompts [B, P] ids; out_ids [B, num_hyp, max_len] (-1 padded), out_lens / out_scores [B, num_hyp], no_speech [B]
int ref_whisper_generate(void* handle, const float* features, int B, int n_mels, int T, const int32_t* prompts, int P,
                         int beam_size, float patience, int num_hyp, float length_penalty, int max_length, int suppress_blank,
                         int suppress_default, int32_t* out_ids, int32_t* out_lens, float* out_scores, float* no_speech) {
  auto* w = static_cast<RefWhisper*>(handle);
  return guarded([&] {
    StorageView f = view_f32(features, {B, n_mels, T});
    std::vector<std::vector<size_t>> pr(B);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < P; ++i) pr[b].push_back(prompts[b * P + i]);
    models::WhisperOptions opt;
    opt.beam_size = beam_size;
    opt.patience = patience;
    opt.num_hypotheses = num_hyp;
    opt.length_penalty = length_penalty;
    opt.max_length = max_length;
    opt.return_scores = true;
    opt.return_no_speech_prob = no_speech != nullptr;
    opt.suppress_blank = suppress_blank != 0;
    opt.suppress_tokens = suppress_default ? std::vector<int>{-1} : std::vector<int>{};
    auto futures = w->pool->generate(f, pr, opt);
    for (int b = 0; b < B; ++b) {
      auto r = futures[b].get();
      if (no_speech) no_speech[b] = r.no_speech_prob;
      for (int h = 0; h < num_hyp; ++h) {
        const bool have = h < static_cast<int>(r.sequences_ids.size());
        const int64_t o = static_cast<int64_t>(b) * num_hyp + h;
        out_lens[o] = have ? static_cast<int32_t>(r.sequences_ids[h].size()) : -1;
        out_scores[o] = have && h < static_cast<int>(r.scores.size()) ? r.scores[h] : 0.f;
        for (int i = 0; i < max_length; ++i)
          out_ids[o * max_length + i] = (have && i < out_lens[o]) ? static_cast<int32_t>(r.sequences_ids[h][i]) : -1;
      }
    }
  });
}

// Full-sequence forward: ids [B,T] -> logits (or log-probs) [B,T,V] fp32.
int ref_forward(void* handle, const int32_t* ids, int B, int T, int return_log_probs,
                float* out, int64_t out_capacity) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<size_t>> v(B, std::vector<size_t>(T));
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < T; ++t)
        v[b][t] = ids[b * T + t];
    StorageView logits = g->generator->forward_batch_async(v, return_log_probs != 0).get();
    StorageView f = logits.to_float32().to(Device::CPU);      // the CUDA build returns device memory
    if (f.size() > out_capacity)
      throw std::runtime_error("ref_forward: output buffer too small");
    std::memcpy(out, f.data<float>(), f.size() * sizeof(float));
  });
}

int ref_vocab_size(void* handle) {
  return (int)static_cast<RefGenerator*>(handle)->vocab->size();
}

void ref_set_num_threads(int n) { set_num_threads(n); }

// ---- op-level entry points (CPU, fp32 activations) ----

// ops::Quantize (src/ops/quantize.cc:21-50): x [rows,cols] -> q int8, scale [rows]
int ref_quantize(const float* x, int rows, int cols, int round_before_cast, int8_t* q, float* scale) {
  return guarded([&] {
    StorageView in = view_f32(x, {rows, cols});
    StorageView out(DataType::INT8), sc(DataType::FLOAT32);
    ops::Quantize(ops::Quantize::ScaleType::GLOBAL, false, round_before_cast != 0)(in, out, sc);
    std::memcpy(q, out.data<int8_t>(), (size_t)rows * cols);
    std::memcpy(scale, sc.data<float>(), rows * sizeof(float));
  });
}

// ops::Gemm int8 (src/ops/gemm.cc:45-107), alpha=1 beta=0 trans_b: a [m,k], b [n,k] -> c [m,n] int32
int ref_gemm_s8(const int8_t* a, const int8_t* b, int m, int n, int k, int32_t* c) {
  return guarded([&] {
    StorageView A({m, k}, const_cast<int8_t*>(a), Device::CPU);
    StorageView B({n, k}, const_cast<int8_t*>(b), Device::CPU);
    StorageView C(DataType::INT32);
    ops::Gemm(1.f, 0.f, false, true)(A, B, C);
    std::memcpy(c, C.data<int32_t>(), (size_t)m * n * sizeof(int32_t));
  });
}

// ops::Gemm fp32 with bias / residual / activation (src/ops/gemm.cc:10-43)
int ref_gemm_f32(const float* a, const float* b, const float* bias, const float* residual, int act,
                 int m, int n, int k, float* c) {
  return guarded([&] {
    StorageView A = view_f32(a, {m, k}), B = view_f32(b, {n, k});
    StorageView C(DataType::FLOAT32);
    StorageView biasv, resv;
    if (bias) biasv = view_f32(bias, {n});
    if (residual) resv = view_f32(residual, {m, n});
    ops::ActivationType at = static_cast<ops::ActivationType>(act < 0 ? 0 : act);
    ops::Gemm(1.f, 0.f, false, true, false, false, act >= 0 ? &at : nullptr)(
      A, B, C, nullptr, bias ? &biasv : nullptr, residual ? &resv : nullptr);
    std::memcpy(c, C.data<float>(), (size_t)m * n * sizeof(float));
  });
}

// ops::Dequantize gemm-output form (src/ops/dequantize.cc:46-59); act<0 = none.
int ref_dequantize_gemm(const int32_t* c, const float* a_scale, const float* b_scale, const float* bias,
                        int act, int m, int n, float* y) {
  return guarded([&] {
    StorageView C({m, n}, const_cast<int32_t*>(c), Device::CPU);
    StorageView sa = view_f32(a_scale, {m}), sb = view_f32(b_scale, {n});
    StorageView biasv;
    if (bias) biasv = view_f32(bias, {n});
    StorageView Y(DataType::FLOAT32);
    ops::ActivationType at = static_cast<ops::ActivationType>(act < 0 ? 0 : act);
    ops::Dequantize(act >= 0 ? &at : nullptr)(C, sa, sb, false, true, Y, bias ? &biasv : nullptr);
    std::memcpy(y, Y.data<float>(), (size_t)m * n * sizeof(float));
  });
}

// ops::RMSNorm (src/ops/rms_norm.cc)
int ref_rms_norm(const float* gamma, const float* x, int rows, int cols, float eps, float* y) {
  return guarded([&] {
    StorageView G = view_f32(gamma, {cols}), X = view_f32(x, {rows, cols});
    StorageView Y(DataType::FLOAT32);
    ops::RMSNorm(eps, false)(G, X, Y);
    std::memcpy(y, Y.data<float>(), (size_t)rows * cols * sizeof(float));
  });
}

// ops::Rotary (src/ops/rotary.cc): x [b,h,t,d] (is_transpose=true), sin/cos [t,ndims]
int ref_rotary(const float* x, const float* sin, const float* cos, int b, int h, int t, int d,
               int ndims, int interleave, float* y) {
  return guarded([&] {
    StorageView X = view_f32(x, {b, h, t, d});
    StorageView S = view_f32(sin, {t, ndims}), C = view_f32(cos, {t, ndims});
    StorageView Y(DataType::FLOAT32);
    ops::Rotary(ndims, interleave != 0)(X, S, C, Y, true);
    std::memcpy(y, Y.data<float>(), (size_t)b * h * t * d * sizeof(float));
  });
}

// ops::SoftMax / LogSoftMax with optional per-row lengths (src/ops/softmax.cc:28-47)
int ref_softmax(const float* x, const int32_t* lengths, int rows, int cols, int log, float* y) {
  return guarded([&] {
    StorageView X = view_f32(x, {rows, cols});
    StorageView L;
    if (lengths) L = StorageView({rows}, const_cast<int32_t*>(lengths), Device::CPU);
    StorageView Y(DataType::FLOAT32);
    ops::SoftMax(log != 0)(X, lengths ? &L : nullptr, Y);
    std::memcpy(y, Y.data<float>(), (size_t)rows * cols * sizeof(float));
  });
}

// ops::TopK (src/ops/topk.cc:14-22)
int ref_topk(const float* x, int rows, int cols, int k, float* values, int32_t* indices) {
  return guarded([&] {
    StorageView X = view_f32(x, {rows, cols});
    StorageView V(DataType::FLOAT32), I(DataType::INT32);
    const ops::TopK topk_op(k);
    topk_op(X, V, I);
    std::memcpy(values, V.data<float>(), (size_t)rows * k * sizeof(float));
    std::memcpy(indices, I.data<int32_t>(), (size_t)rows * k * sizeof(int32_t));
  });
}

// ops::Gather axis 0 (src/ops/gather.cc:49-86): data [n,d], ids [m] -> [m,d]
int ref_gather(const float* data, int n, int d, const int32_t* ids, int m, float* out) {
  return guarded([&] {
    StorageView D = view_f32(data, {n, d});
    StorageView I({m}, const_cast<int32_t*>(ids), Device::CPU);
    StorageView O(DataType::FLOAT32);
    ops::Gather(0, 0)(D, I, O);
    std::memcpy(out, O.data<float>(), (size_t)m * d * sizeof(float));
  });
}

}  // extern "C"

// layers::RotaryEmbeddings (src/layers/attention_layer.cc:178-343) applied to x [1,1,T,dim] at `offset`:
// lets the tests recover the reference's own sin/cos tables (incl. Linear / Llama3 frequency scaling).
#include <ctranslate2/layers/attention_layer.h>
extern "C" int ref_rotary_embeddings(const float* x, int t, int dim, int offset, int interleave, int scaling_type,
                                     float scaling_factor, float base, float low_freq_factor, float high_freq_factor,
                                     int original_max_position_embeddings, float* y) {
  return guarded([&] {
    layers::RotaryEmbeddings rot(dim, interleave != 0, static_cast<layers::RotaryScalingType>(scaling_type),
                                 scaling_factor, base, /*num_initial_positions=*/2048, nullptr, nullptr,
                                 low_freq_factor, high_freq_factor, original_max_position_embeddings, 0, true);
    StorageView X({1, 1, t, dim}, std::vector<float>(x, x + static_cast<size_t>(t) * dim), Device::CPU);
    rot.apply(X, offset);
    std::memcpy(y, X.data<float>(), static_cast<size_t>(t) * dim * sizeof(float));
  });
}

// ops::LayerNorm (src/ops/layer_norm.cc) over the last axis
extern "C" int ref_layer_norm(const float* gamma, const float* beta, const float* x, int rows, int cols, float eps,
                              float* y) {
  return guarded([&] {
    StorageView G = view_f32(gamma, {cols}), Bt = view_f32(beta, {cols}), X = view_f32(x, {rows, cols});
    StorageView Y(DataType::FLOAT32);
    ops::LayerNorm(-1, eps)(Bt, G, X, Y);
    std::memcpy(y, Y.data<float>(), (size_t)rows * cols * sizeof(float));
  });
}

// layers::TransformerEncoder of the translator's model: ids [B,S] (padded with any id), lengths [B] -> out [B,S,d] fp32
#include <ctranslate2/layers/transformer.h>
extern "C" int ref_encoder_forward(void* handle, const int32_t* ids, const int32_t* lengths, int B, int S, float* out,
                                   int64_t out_capacity) {
  auto* t = static_cast<RefTranslator*>(handle);
  return guarded([&] {
    layers::TransformerEncoder encoder(*t->model, "encoder");
    StorageView I({B, S}, DataType::INT32), L({B}, DataType::INT32);
    std::memcpy(I.data<int32_t>(), ids, sizeof(int32_t) * B * S);
    std::memcpy(L.data<int32_t>(), lengths, sizeof(int32_t) * B);
    StorageView O(DataType::FLOAT32);
    encoder({I}, &L, O);
    if (O.size() > out_capacity)
      throw std::runtime_error("ref_encoder_forward: output buffer too small");
    std::memcpy(out, O.data<float>(), sizeof(float) * O.size());
  });
}

// Greedy generate_batch with the logits processors of GenerationOptions (decoding.cc:1099-1112, decoding_utils.cc):
// repetition_penalty, no_repeat_ngram_size, disable_unk, suppress_sequences.  `suppress` is a flat list of token ids in which
// -1 ends a sequence (n_suppress entries).  Scores are returned too (return_scores = true, length_penalty 1).
extern "C" int ref_generate_processors(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int min_len,
                                       int end_id, float repetition_penalty, int no_repeat_ngram_size, int disable_unk,
                                       const int32_t* suppress, int n_suppress, int32_t* out_ids, int32_t* out_lens,
                                       float* out_scores) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = min_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = true;
    opt.length_penalty = 1.f;
    opt.repetition_penalty = repetition_penalty;
    opt.no_repeat_ngram_size = static_cast<size_t>(no_repeat_ngram_size);
    opt.disable_unk = disable_unk != 0;
    std::vector<std::string> current;
    for (int i = 0; i < n_suppress; ++i) {
      if (suppress[i] < 0) {
        if (!current.empty())
          opt.suppress_sequences.push_back(std::move(current));
        current.clear();
      } else {
        current.push_back(g->vocab->to_token(suppress[i]));
      }
    }
    if (!current.empty())
      opt.suppress_sequences.push_back(std::move(current));
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    auto futures = g->generator->generate_batch_async(prompts, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      const auto& ids = result.sequences_ids.at(0);
      out_lens[b] = static_cast<int32_t>(ids.size());
      out_scores[b] = result.scores.at(0);
      for (int t = 0; t < max_len; ++t)
        out_ids[b * max_len + t] = t < (int)ids.size() ? (int32_t)ids[t] : -1;
    }
  });
}

// Greedy generate_batch over prompts of DIFFERENT lengths: prompt_ids [B,P] right-padded with -1.  return_scores = true.
extern "C" int ref_generate_ragged(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int min_len,
                                   int end_id, int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P && prompt_ids[b * P + t] >= 0; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = min_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = true;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    auto futures = g->generator->generate_batch_async(prompts, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      const auto& ids = result.sequences_ids.at(0);
      out_lens[b] = static_cast<int32_t>(ids.size());
      out_scores[b] = result.scores.at(0);
      for (int t = 0; t < max_len + P; ++t)
        out_ids[b * (max_len + P) + t] = t < (int)ids.size() ? (int32_t)ids[t] : -1;
    }
  });
}

// Generator::score_batch (ScoringOptions defaults except `offset`): ids [B,P] right-padded with -1; out_scores [B, P-1] holds
// the log-probability of ids[b][t+1] given ids[b][:t+1] for t + 1 < length (0 elsewhere); out_lens[b] = number of scores.
#include <ctranslate2/scoring.h>
extern "C" int ref_score(void* handle, const int32_t* ids, int B, int P, int offset, float* out_scores, int32_t* out_lens) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> tokens(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P && ids[b * P + t] >= 0; ++t)
        tokens[b].push_back(g->vocab->to_token(ids[b * P + t]));
    ScoringOptions opt;
    opt.offset = offset;
    auto futures = g->generator->score_batch_async(tokens, opt);
    for (int b = 0; b < B; ++b) {
      auto result = futures[b].get();
      out_lens[b] = static_cast<int32_t>(result.tokens_score.size());
      for (int t = 0; t < P - 1; ++t)
        out_scores[b * (P - 1) + t] = t < out_lens[b] ? result.tokens_score[t] : 0.f;
    }
  });
}

#ifdef REF_DRIVER_CUDA
// ---- the reference's CUDA specialisations, run on device buffers staged from host arrays ------------------------------
// (GPU-side oracle: tools/make_golden_cuda.py turns their outputs into fixtures under tests/golden/, tests compare live
// when the library is present on the GPU box.)
#include <ctranslate2/ops/awq/gemm.h>
#include <ctranslate2/ops/awq/gemv.h>
#include <ctranslate2/ops/awq/dequantize_awq.h>
#include <cuda_runtime.h>
#include <chrono>

namespace {
  StorageView to_cuda(const void* host, Shape shape, DataType dt) {
    StorageView h(dt, Device::CPU);
    h.resize(shape);
    std::memcpy(h.buffer(), host, h.size() * h.item_size());
    return h.to(Device::CUDA);
  }
  void to_host(const StorageView& d, void* out) {
    StorageView h = d.to(Device::CPU);
    std::memcpy(out, h.buffer(), h.size() * h.item_size());
  }
}

extern "C" {

// ops::GemmAwq (src/ops/awq/gemm.cc:8-33): x f16 [m,k], qweight int32 [k, n/8], scales f16 [k/g, n], qzeros int32 [k/g, n/8]
// -> y f16 [m, n]
int ref_cuda_gemm_awq(const uint16_t* x, const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros, int m, int n,
                      int k, int group, uint16_t* y) {
  return guarded([&] {
    StorageView X = to_cuda(x, {m, k}, DataType::FLOAT16);
    StorageView W = to_cuda(qweight, {k, n / 8}, DataType::INT32);
    StorageView S = to_cuda(scales, {k / group, n}, DataType::FLOAT16);
    StorageView Z = to_cuda(qzeros, {k / group, n / 8}, DataType::INT32);
    StorageView Y(DataType::FLOAT16, Device::CUDA);
    ops::GemmAwq op(1.f, 0.f, false, false, false, false, nullptr);
    op(X, W, S, Z, Y);
    to_host(Y, y);
  });
}

// ops::GemvAwq (src/ops/awq/gemv.cc:9-37; m <= 8: gemv kernel, else gemv2): x f16 [m,k], qweight int32 [n, k/8],
// scales f16 [n, sw], qzeros int32 [n, zw] (padded widths of the AWQ_GEMV layout) -> y f16 [m, n]
int ref_cuda_gemv_awq(const uint16_t* x, const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros, int m, int n,
                      int k, int sw, int zw, uint16_t* y) {
  return guarded([&] {
    StorageView X = to_cuda(x, {m, k}, DataType::FLOAT16);
    StorageView W = to_cuda(qweight, {n, k / 8}, DataType::INT32);
    StorageView S = to_cuda(scales, {n, sw}, DataType::FLOAT16);
    StorageView Z = to_cuda(qzeros, {n, zw}, DataType::INT32);
    StorageView Y(DataType::FLOAT16, Device::CUDA);
    ops::GemvAwq op(1.f, 0.f, false, false, false, false, nullptr);
    op(X, W, S, Z, Y);
    to_host(Y, y);
  });
}

// ops::DequantizeAwq (src/ops/awq/dequantize.cc): qweight int32 [k, n/8] (AWQ_GEMM layout) -> f16 [k, n]
int ref_cuda_dequantize_awq(const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros, int n, int k, int group,
                            uint16_t* w) {
  return guarded([&] {
    StorageView W = to_cuda(qweight, {k, n / 8}, DataType::INT32);
    StorageView S = to_cuda(scales, {k / group, n}, DataType::FLOAT16);
    StorageView Z = to_cuda(qzeros, {k / group, n / 8}, DataType::INT32);
    StorageView Y(DataType::FLOAT16, Device::CUDA);
    ops::DequantizeAwq()(W, S, Z, Y);
    to_host(Y, w);
  });
}

// layers::Dense INT8 arm on the GPU (common.cc:353-401): Quantize -> cublasGemmEx s8 -> Dequantize(+bias, act), x f16 [m,k]
int ref_cuda_dense_s8(const uint16_t* x, const int8_t* w, const float* w_scale, int act, int m, int n, int k, uint16_t* y) {
  return guarded([&] {
    StorageView X = to_cuda(x, {m, k}, DataType::FLOAT16);
    StorageView W = to_cuda(w, {n, k}, DataType::INT8);
    StorageView WS = to_cuda(w_scale, {n}, DataType::FLOAT32);
    StorageView Q(DataType::INT8, Device::CUDA), QS(DataType::FLOAT32, Device::CUDA), C(DataType::INT32, Device::CUDA),
        Y(DataType::FLOAT16, Device::CUDA);
    ops::Quantize(ops::Quantize::ScaleType::GLOBAL, false, true)(X, Q, QS);
    ops::Gemm(1.f, 0.f, false, true)(Q, W, C);
    ops::ActivationType at = static_cast<ops::ActivationType>(act < 0 ? 0 : act);
    ops::Dequantize(act >= 0 ? &at : nullptr)(C, QS, WS, false, true, Y, nullptr);
    to_host(Y, y);
  });
}

// Wall time (seconds) of one greedy generate_batch of exactly max_len tokens per row on the current device, after a
// cudaDeviceSynchronize on both sides; out_ids as ref_generate.
int ref_generate_timed(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int end_id, int32_t* out_ids,
                       double* seconds) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = max_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = false;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    auto futures = g->generator->generate_batch_async(prompts, opt, /*max_batch_size=*/0);
    std::vector<GenerationResult> results;
    for (auto& f : futures) results.push_back(f.get());
    cudaDeviceSynchronize();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out_ids)
      for (int b = 0; b < B; ++b) {
        const auto& ids = results[b].sequences_ids.at(0);
        for (int t = 0; t < max_len; ++t)
          out_ids[b * max_len + t] = t < (int)ids.size() ? (int32_t)ids[t] : -1;
      }
  });
}

// Per-step host timestamps of one greedy generate_batch through the public per-step callback (generation.h:77): stamps[s] =
// seconds since the call started at which step s of batch entry 0 was delivered.  The prompt pass ends at stamps[0]; the decode
// time per step is the slope of the later stamps — no difference of two prompt-dominated wall times is needed.
int ref_generate_steps(void* handle, const int32_t* prompt_ids, int B, int P, int max_len, int end_id, double* stamps,
                       double* seconds) {
  auto* g = static_cast<RefGenerator*>(handle);
  return guarded([&] {
    std::vector<std::vector<std::string>> prompts(B);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < P; ++t)
        prompts[b].push_back(g->vocab->to_token(prompt_ids[b * P + t]));
    GenerationOptions opt;
    opt.beam_size = 1;
    opt.sampling_topk = 1;
    opt.max_length = max_len;
    opt.min_length = max_len;
    opt.include_prompt_in_result = false;
    opt.return_scores = false;
    opt.end_token = std::vector<size_t>{static_cast<size_t>(end_id)};
    for (int s = 0; s < max_len; ++s) stamps[s] = -1.0;
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    opt.callback = [&](GenerationStepResult r) {
      if (r.batch_id == 0 && r.step < static_cast<size_t>(max_len) && stamps[r.step] < 0)
        stamps[r.step] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      return false;
    };
    auto futures = g->generator->generate_batch_async(prompts, opt, /*max_batch_size=*/0);
    for (auto& f : futures) f.get();
    cudaDeviceSynchronize();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  });
}

}  // extern "C"
#endif  // REF_DRIVER_CUDA

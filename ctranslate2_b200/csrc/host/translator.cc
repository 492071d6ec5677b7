// translator.cc — see translator.h.  Reference counterparts: src/models/transformer.cc, src/models/sequence_to_sequence.cc
// (EncoderDecoderReplica::run_translation), src/layers/transformer.cc (TransformerEncoder / TransformerDecoder),
// src/layers/attention.cc, src/layers/common.cc (Embeddings, position encoders, LayerNorm, Dense), src/decoding.cc (BeamSearch).
#include "translator.h"

#include <cuda_profiler_api.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace ct2b200 {

// =============================================================================================
// configuration
// =============================================================================================
namespace {

const HostVariable* find_any(const ModelFile& f, std::initializer_list<std::string> names) {
  for (const auto& n : names)
    if (const HostVariable* v = f.find(n)) return v;
  return nullptr;
}

std::string embeddings_scope(const ModelFile& f, const std::string& scope) {
  if (f.find(scope + "/embeddings_0/weight")) {
    if (f.find(scope + "/embeddings_1/weight"))
      throw std::invalid_argument("models with several input features (embeddings_1) are not supported");
    return scope + "/embeddings_0";
  }
  return scope + "/embeddings";
}

// build_embeddings_scale (transformer.cc:380-402): the attribute is a flag or the scale itself
float embeddings_scale(const ModelFile& f, const std::string& scope, int64_t depth) {
  const HostVariable* s = find_any(f, {scope + "/scale_embeddings", scope + "/embeddings/multiply_by_sqrt_depth"});
  if (!s || (s->type_id == 1 && s->scalar() != 0.0)) return std::sqrt(static_cast<float>(depth));
  if (s->type_id != 1 && s->scalar() != 1.0) return static_cast<float>(s->scalar());
  return 0.f;
}

double scoped_attribute(const ModelFile& f, const std::string& scope, const std::string& name, double fallback) {
  // spec revisions < 5 keep the attribute at the top level (models/transformer.cc:67-79)
  const HostVariable* v = find_any(f, {scope + "/" + name, name});
  return v ? v->scalar() : fallback;
}

}  // namespace

Seq2SeqConfig parse_seq2seq_config(const ModelFile& f) {
  Seq2SeqConfig mc;
  if (!f.find("encoder/layer_0/self_attention/linear_0/weight") || !f.find("decoder/layer_0/attention/linear_0/weight"))
    throw std::invalid_argument("ct2b200 Translator serves encoder-decoder Transformer models (TransformerSpec, WhisperSpec); got " +
                                f.spec_name);
  while (f.find("encoder/layer_" + std::to_string(mc.enc_layers) + "/self_attention/linear_0/weight")) ++mc.enc_layers;
  while (f.find("decoder/layer_" + std::to_string(mc.dec_layers) + "/self_attention/linear_0/weight")) ++mc.dec_layers;
  const HostVariable& temb = f.get("decoder/embeddings/weight");
  mc.tgt_vocab = temb.shape[0];
  mc.d_model = temb.shape[1];
  mc.whisper = f.find("encoder/conv1/weight") != nullptr;
  if (mc.whisper) {
    // WhisperEncoder (layers/whisper.cc:8-23): Conv1D(k 3, stride 1, pad 1) + GELU, Conv1D(k 3, stride 2, pad 1) + GELU,
    // stored positions, pre-norm GELU layers
    const HostVariable& c1 = f.get("encoder/conv1/weight");
    const HostVariable& c2 = f.get("encoder/conv2/weight");
    CT2_REQUIRE(c1.shape.size() == 3 && c1.shape[2] == 3 && c2.shape.size() == 3 && c2.shape[2] == 3, "Whisper convolutions must have kernel size 3");
    CT2_REQUIRE(c1.shape[0] == mc.d_model && c2.shape[0] == mc.d_model && c2.shape[1] == mc.d_model, "unexpected convolution shapes");
    mc.n_mels = c1.shape[1];
    mc.max_frames = f.get("encoder/position_encodings/encodings").shape[0];
    CT2_REQUIRE(f.find("decoder/position_encodings/encodings") != nullptr, "Whisper decoders store their position encodings");
  } else {
    const HostVariable& semb = f.get(embeddings_scope(f, "encoder") + "/weight");
    mc.src_vocab = semb.shape[0];
    CT2_REQUIRE(semb.shape[1] == mc.d_model, "encoder and decoder depths differ");
  }
  // num_heads: attribute since revision 3; TransformerBase / TransformerBig imply 8 / 16 before (models/transformer.cc:62-65)
  mc.num_heads = static_cast<int>(scoped_attribute(f, "encoder", "num_heads", f.spec_name == "TransformerBig" ? 16 : 8));
  CT2_REQUIRE(static_cast<int>(scoped_attribute(f, "decoder", "num_heads", mc.num_heads)) == mc.num_heads,
              "encoder and decoder head counts differ");
  CT2_REQUIRE(mc.d_model % mc.num_heads == 0, "d_model must be divisible by num_heads");
  mc.head_dim = static_cast<int>(mc.d_model / mc.num_heads);
  mc.enc_pre_norm = scoped_attribute(f, "encoder", "pre_norm", 1.0) != 0.0;
  mc.dec_pre_norm = scoped_attribute(f, "decoder", "pre_norm", 1.0) != 0.0;
  mc.enc_activation = static_cast<int>(scoped_attribute(f, "encoder", "activation", 0.0));
  mc.dec_activation = static_cast<int>(scoped_attribute(f, "decoder", "activation", 0.0));
  if (mc.whisper) {
    mc.enc_pre_norm = true;
    mc.enc_activation = CT2B200_ACT_GELU;
  }
  mc.enc_emb_scale = mc.whisper ? 0.f : embeddings_scale(f, "encoder", mc.d_model);
  mc.dec_emb_scale = embeddings_scale(f, "decoder", mc.d_model);
  mc.ffn_dim = f.get("encoder/layer_0/ffn/linear_0/weight").shape[0];
  mc.round_before_cast = f.binary_version >= 5;
  mc.has_enc_final_norm = f.find("encoder/layer_norm/gamma") != nullptr;
  mc.has_dec_final_norm = f.find("decoder/layer_norm/gamma") != nullptr;
  const bool has_beta = f.find("encoder/layer_0/self_attention/layer_norm/beta") != nullptr;
  mc.eps = static_cast<float>(f.config_number("layer_norm_epsilon", has_beta ? 1e-5 : 1e-6));
  // what TransformerSpec can express and this engine does not compute: refuse instead of translating something else
  auto absent = [&](const std::string& name, const char* what) {
    if (f.find(name)) throw std::invalid_argument(std::string(what) + " (" + name + ") is not supported by the Translator engine");
  };
  CT2_REQUIRE(has_beta, "RMSNorm encoder-decoder models are not supported (LayerNorm with beta is)");
  for (const char* scope : {"encoder", "decoder"}) {
    const std::string s(scope);
    absent(s + "/layernorm_embedding/gamma", "layernorm_embedding");
    absent(s + "/layer_0/self_attention/relative_position_keys", "relative position representations");
    absent(s + "/layer_0/self_attention/relative_attention_bias", "relative attention bias");
    absent(s + "/layer_0/self_attention/rotary_dim", "rotary embeddings");
    absent(s + "/layer_0/self_attention/num_heads_kv", "grouped-query attention");
    absent(s + "/layer_0/ffn/linear_0_noact/weight", "gated feed-forward layers");
    absent(s + "/project_in/weight", "project_in");
    absent(s + "/project_out/weight", "project_out");
    if (scoped_attribute(f, s, "alibi", 0.0) != 0.0) throw std::invalid_argument("ALiBi is not supported by the Translator engine");
    if (scoped_attribute(f, s, "embeddings_merge", 0.0) != 0.0)
      throw std::invalid_argument("embeddings_merge other than CONCAT of one feature is not supported");
  }
  mc.start_from_zero_embedding = f.attribute("decoder/start_from_zero_embedding", 0.0) != 0.0;
  absent("decoder/scale_outputs", "scaled outputs");
  absent("decoder/layer_0/layer_scalar", "layer_scalar");
  absent("decoder/layer_0/self_attention/queries_scale", "a custom queries_scale");
  if (f.attribute("decoder/final_logit_softcapping", 0.0) != 0.0)
    throw std::invalid_argument("final logit soft-capping is not supported by the Translator engine");
  CT2_REQUIRE(f.revision != 1, "spec revision 1 (OpenNMT-tf variable names) is not supported");
  const HostVariable& w = f.get("encoder/layer_0/self_attention/linear_0/weight");
  mc.weights = w.type_id == 1 ? "int8" : w.type_id == 2 ? "int16" : w.type_id == 4 ? "float16" : w.type_id == 5 ? "bfloat16" : "float32";
  CT2_REQUIRE(w.type_id != 2, "int16 models are not supported (convert with int8 or a float type)");
  return mc;
}

// =============================================================================================
// loading
// =============================================================================================
void Translator::load_dense(const ModelFile& f, const std::string& prefix, DenseWeights& w) {
  const bool int8 = load_dense_matrix(f, prefix, dtype_, weight_type_, stream_, w.weight, w.scale, w.n, w.k);
  w.kind = int8 ? DenseWeights::INT8 : DenseWeights::FLOAT16;
  CT2_REQUIRE(!int8 || w.k % 16 == 0, "int8 Dense layers need an input size that is a multiple of 16");
  mc_.weight_bytes += w.weight.bytes + w.scale.bytes;
  if (const HostVariable* b = f.find(prefix + "/bias")) {
    const auto bytes = convert_to_dtype(*b, dtype_);
    upload(w.bias, bytes.data(), bytes.size());
  }
}

void Translator::load_norm(const ModelFile& f, const std::string& prefix, NormWeights& n) {
  const auto g = convert_to_dtype(f.get(prefix + "/gamma"), dtype_);
  const auto b = convert_to_dtype(f.get(prefix + "/beta"), dtype_);
  upload(n.gamma, g.data(), g.size());
  upload(n.beta, b.data(), b.size());
}

namespace {
// generate_sinusoidal_position_encoding (common.cc:204-229): positions start at 1, [sin | cos] halves, fp32 then cast
std::vector<float> sinusoidal_positions(int64_t max_time, int64_t depth) {
  const float inc = std::log(10000.f) / static_cast<float>(depth / 2 - 1);
  std::vector<float> ts(depth / 2);
  for (int64_t i = 0; i < depth / 2; ++i) ts[i] = std::exp(-inc * static_cast<float>(i));
  std::vector<float> e(max_time * depth);
  for (int64_t t = 0; t < max_time; ++t)
    for (int64_t j = 0; j < depth / 2; ++j) {
      const float a = static_cast<float>(t + 1) * ts[j];
      e[t * depth + j] = std::sin(a);
      e[t * depth + depth / 2 + j] = std::cos(a);
    }
  return e;
}
}  // namespace

Translator::Translator(const std::string& model_dir, const ct2b200_generator_config& cfg) {
  device_ = cfg.device;
  CT2_CUDA_CHECK(cudaSetDevice(device_));
  int major = 0;
  CT2_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device_));
  if (major != 10)
    throw std::runtime_error("ct2b200 needs an sm_100 (B200) device; found compute capability major " + std::to_string(major));
  CT2_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
  CT2_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  dtype_ = cfg.compute_type;
  weight_type_ = cfg.weight_type;
  use_graph_ = cfg.use_cuda_graph != 0;
  CT2_REQUIRE(cfg.tp_size <= 1, "the Translator engine does not run tensor parallel");

  ModelFile f(model_dir);
  mc_ = parse_seq2seq_config(f);
  if (mc_.whisper) {
    // the convolutions run as im2col + float Dense: weights [d, Cin, 3] flattened to [d, Cin * 3] in T (the reference keeps
    // them in float on CUDA too, model.cc:204-223; an int8-stored convolution is dequantized here: w = q / scale)
    auto load_conv = [&](const std::string& prefix, DenseWeights& w) {
      const HostVariable& wt = f.get(prefix + "/weight");
      HostVariable flat = wt;
      std::vector<float> deq;
      if (wt.type_id == 1) {
        const HostVariable& sc = f.get(prefix + "/weight_scale");
        CT2_REQUIRE(sc.type_id == 0 && (sc.size() == wt.shape[0] || sc.size() == 1), "unexpected convolution weight_scale");
        deq.resize(wt.size());
        const int64_t per = wt.size() / wt.shape[0];
        for (int64_t i = 0; i < wt.size(); ++i) {
          float scale;
          std::memcpy(&scale, sc.data + 4 * (sc.size() == 1 ? 0 : i / per), 4);
          deq[i] = static_cast<float>(reinterpret_cast<const int8_t*>(wt.data)[i]) / scale;
        }
        flat.type_id = 0;
        flat.data = reinterpret_cast<const uint8_t*>(deq.data());
        flat.nbytes = deq.size() * 4;
      }
      const auto bytes = convert_to_dtype(flat, dtype_);
      upload(w.weight, bytes.data(), bytes.size());
      w.kind = DenseWeights::FLOAT16;
      w.n = wt.shape[0];
      w.k = wt.shape[1] * wt.shape[2];
      mc_.weight_bytes += bytes.size();
      if (const HostVariable* b = f.find(prefix + "/bias")) {
        const auto bb = convert_to_dtype(*b, dtype_);
        upload(w.bias, bb.data(), bb.size());
      }
    };
    CT2_REQUIRE(dtype_ == CT2B200_F32 || (mc_.n_mels * 3) % 8 == 0, "n_mels * 3 must be a multiple of 8 for float16 / bfloat16");
    load_conv("encoder/conv1", conv1_);
    load_conv("encoder/conv2", conv2_);
  } else {
    load_dense(f, embeddings_scope(f, "encoder"), enc_emb_);
  }
  load_dense(f, "decoder/embeddings", dec_emb_);
  load_dense(f, "decoder/projection", projection_);
  if (mc_.has_enc_final_norm) load_norm(f, "encoder/layer_norm", enc_norm_);
  if (mc_.has_dec_final_norm) load_norm(f, "decoder/layer_norm", dec_norm_);
  enc_.resize(mc_.enc_layers);
  for (int l = 0; l < mc_.enc_layers; ++l) {
    const std::string p = "encoder/layer_" + std::to_string(l) + "/";
    load_norm(f, p + "self_attention/layer_norm", enc_[l].self.norm);
    load_dense(f, p + "self_attention/linear_0", enc_[l].self.in);
    load_dense(f, p + "self_attention/linear_1", enc_[l].self.out);
    load_norm(f, p + "ffn/layer_norm", enc_[l].ffn.norm);
    load_dense(f, p + "ffn/linear_0", enc_[l].ffn.ff1);
    load_dense(f, p + "ffn/linear_1", enc_[l].ffn.ff2);
  }
  dec_.resize(mc_.dec_layers);
  for (int l = 0; l < mc_.dec_layers; ++l) {
    const std::string p = "decoder/layer_" + std::to_string(l) + "/";
    load_norm(f, p + "self_attention/layer_norm", dec_[l].self.norm);
    load_dense(f, p + "self_attention/linear_0", dec_[l].self.in);
    load_dense(f, p + "self_attention/linear_1", dec_[l].self.out);
    load_norm(f, p + "attention/layer_norm", dec_[l].cross.norm);
    load_dense(f, p + "attention/linear_0", dec_[l].cross.in);
    load_dense(f, p + "attention/linear_1", dec_[l].cross.kv);
    load_dense(f, p + "attention/linear_2", dec_[l].cross.out);
    load_norm(f, p + "ffn/layer_norm", dec_[l].ffn.norm);
    load_dense(f, p + "ffn/linear_0", dec_[l].ffn.ff1);
    load_dense(f, p + "ffn/linear_1", dec_[l].ffn.ff2);
  }
  // position encodings: stored table (PositionEmbedding) or sinusoidal (SinusoidalPositionEncoder, 500 positions or more)
  auto load_positions = [&](const std::string& scope, DeviceBuffer& dst) -> int64_t {
    if (const HostVariable* e = f.find(scope + "/position_encodings/encodings")) {
      const auto bytes = convert_to_dtype(*e, dtype_);
      upload(dst, bytes.data(), bytes.size());
      return e->shape[0];
    }
    const int64_t count = std::max<int64_t>(500, cfg.max_length);
    const std::vector<float> enc = sinusoidal_positions(count, mc_.d_model);
    HostVariable v;
    v.shape = {count, mc_.d_model};
    v.type_id = 0;
    v.data = reinterpret_cast<const uint8_t*>(enc.data());
    v.nbytes = enc.size() * 4;
    const auto bytes = convert_to_dtype(v, dtype_);
    upload(dst, bytes.data(), bytes.size());
    return count;
  };
  enc_positions_ = load_positions("encoder", enc_pos_);
  dec_positions_ = load_positions("decoder", dec_pos_);
  SplitKWorkspace::get(stream_);   // create the split-K scratch outside any graph capture
  CT2_CUDA_CHECK(cudaDeviceSynchronize());
}

Translator::~Translator() {
  if (graph_) cudaGraphExecDestroy(graph_);
  if (host_pinned_) cudaFreeHost(host_pinned_);
  if (stream_) {
    cudaStreamSynchronize(stream_);
    SplitKWorkspace::release(stream_);
    cudaStreamDestroy(stream_);
  }
}

void Translator::ensure_arena(int64_t batch, int64_t src_len, int beam, int64_t max_steps) {
  CT2_REQUIRE(src_len <= enc_positions_ && max_steps <= dec_positions_,
              "No position encodings are defined for positions this far (common.cc:157-161)");
  if (batch <= cap_batch_ && src_len <= cap_src_ && beam <= cap_beam_ && max_steps <= cap_steps_) return;
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (graph_) {
    cudaGraphExecDestroy(graph_);
    graph_ = nullptr;
    graph_key_.clear();
  }
  cap_batch_ = std::max(cap_batch_, batch);
  cap_src_ = std::max(cap_src_, src_len);
  cap_beam_ = std::max(cap_beam_, beam);
  cap_steps_ = std::max(cap_steps_, max_steps);
  const size_t es = dtype_size(dtype_);
  beam_.ensure(cap_batch_, cap_beam_, cap_steps_, es);          // same capacities: its stride is the K/V stride
  const int64_t B = cap_batch_, S = cap_src_, L = cap_steps_, N = B * cap_beam_;
  const int64_t R = std::max(B * S, N), d = mc_.d_model, F = mc_.ffn_dim, V = mc_.tgt_vocab;
  cap_rows_ = R;
  src_ids_.alloc(B * S * 4);
  src_lens_.alloc(B * 4);
  x_.alloc(R * d * es);
  xn_.alloc(R * d * es);
  xq_.alloc(R * std::max(d, F));
  xs_.alloc(R * 4);
  qkv_.alloc(R * 3 * d * es);
  ctx_.alloc(R * d * es);
  h_.alloc(R * F * es);
  q_.alloc(R * d * es);
  memory_.alloc(B * S * d * es);
  mem_kv_.resize(mc_.dec_layers);
  self_k_.resize(mc_.dec_layers);
  self_v_.resize(mc_.dec_layers);
  for (int l = 0; l < mc_.dec_layers; ++l) {
    mem_kv_[l].alloc(B * S * 2 * d * es);
    self_k_[l].alloc(N * L * d * es);
    self_v_[l].alloc(N * L * d * es);
  }
  logits_.alloc(N * ((V + 7) / 8 * 8) * es);
  if (mc_.whisper) {
    const int64_t frames = 2 * S;                        // conv2 halves the frames
    features_.alloc(B * mc_.n_mels * frames * 4);
    cols_.alloc(std::max(B * frames * mc_.n_mels * 3, B * S * d * 3) * es);
    conv_out_.alloc(B * frames * d * es);
  }
  const size_t need = static_cast<size_t>(B) * S + B + 64 + static_cast<size_t>(N) * 16 + 8192;
  if (need > host_pinned_elems_) {
    if (host_pinned_) cudaFreeHost(host_pinned_);
    CT2_CUDA_CHECK(cudaMallocHost(&host_pinned_, need * sizeof(int32_t)));
    host_pinned_elems_ = need;
  }
}

// =============================================================================================
// layers
// =============================================================================================
void Translator::dense(const DenseWeights& w, const NormWeights* pre, const void* x, int64_t rows, const void* residual,
                       int act, void* y, bool prequantized, int64_t ldy) {
  if (ldy == 0) ldy = w.n;
  if (w.kind == DenseWeights::INT8) {
    if (prequantized)
      ;                                            // the post-norm kernel before this call left Quantize(x) in xq_ / xs_
    else if (pre)
      launch_layer_norm(x, pre->gamma.ptr, pre->beta.ptr, rows, w.k, mc_.eps, nullptr, xq_.as<int8_t>(), xs_.as<float>(),
                        mc_.round_before_cast, dtype_, stream_);
    else
      launch_quantize_rows(x, dtype_, rows, w.k, mc_.round_before_cast, xq_.as<int8_t>(), xs_.as<float>(), stream_);
    DenseEpilogue e{xs_.as<float>(), w.scale.as<float>(), w.bias.ptr, residual, y, nullptr, act, ldy};
    gemm_s8(xq_.as<int8_t>(), w.weight.as<int8_t>(), rows, w.n, w.k, e, dtype_, CT2B200_GEMM_AUTO, stream_);
  } else {
    const void* src = x;
    if (pre) {
      launch_layer_norm(x, pre->gamma.ptr, pre->beta.ptr, rows, w.k, mc_.eps, xn_.ptr, nullptr, nullptr, true, dtype_, stream_);
      src = xn_.ptr;
    }
    CT2_REQUIRE(ldy == w.n, "float Dense writes contiguous rows");
    gemm_float(src, w.weight.ptr, w.bias.ptr, residual, act, rows, w.n, w.k, y, dtype_, stream_);
  }
}

// Row stride of the logits for a search: INT8 projections write rows padded to a multiple of 8 elements (16-byte stores in the
// GEMM epilogue and 16-byte loads in the scoring kernel whatever the vocabulary size, e.g. 58101); the beam > 8 path runs
// ops::TopK over the flattened [beam * vocab] scores and needs them contiguous.
void Translator::set_logits_ld(BeamState& bs) {
  const int64_t V = mc_.tgt_vocab;
  logits_ld_ = (projection_.kind == DenseWeights::INT8 && bs.beam <= 8) ? (V + 7) / 8 * 8 : V;
  bs.vocab_ld = logits_ld_;
}

// Post-norm sublayer end (transformer.cc:35-38, attention.cc:603-606): x = LayerNorm(x).  When the consumer is an INT8 Dense
// the same launch also leaves its Quantize (of the rounded output, as the reference's separate op reads it) in xq_ / xs_;
// returns whether it did, for the `prequantized` argument of that Dense.
bool Translator::post_norm(const NormWeights& n, void* x, int64_t rows, const DenseWeights* next) {
  const bool q = next && next->kind == DenseWeights::INT8 && next->k == mc_.d_model;
  launch_layer_norm(x, n.gamma.ptr, n.beta.ptr, rows, mc_.d_model, mc_.eps, x, q ? xq_.as<int8_t>() : nullptr,
                    q ? xs_.as<float>() : nullptr, q ? mc_.round_before_cast : true, dtype_, stream_);
  return q;
}

// TransformerEncoder::operator() (transformer.cc:427-471); rows = batch * S, padded positions are computed and ignored
void Translator::run_encoder(int64_t batch, int64_t S) {
  launch_embed_pos(enc_emb_.weight.ptr, enc_emb_.kind == DenseWeights::INT8 ? enc_emb_.scale.as<float>() : nullptr,
                   src_ids_.as<int32_t>(), batch * S, mc_.d_model, mc_.enc_emb_scale, enc_pos_.ptr, S, nullptr, false, x_.ptr,
                   dtype_, stream_);
  run_encoder_layers(batch, S, src_lens_.as<int32_t>());
}

// the layer stack on x_ -> memory_ (lens_d = null: every position is valid)
void Translator::run_encoder_layers(int64_t batch, int64_t S, const int32_t* lens_d) {
  const int64_t rows = batch * S, d = mc_.d_model;
  const float scale = 1.f / std::sqrt(static_cast<float>(mc_.head_dim));
  const bool pre = mc_.enc_pre_norm;
  bool xq = false;                                 // xq_ / xs_ already hold Quantize(x_) (left by a post-norm launch)
  for (int l = 0; l < mc_.enc_layers; ++l) {
    EncoderLayerWeights& w = enc_[l];
    dense(w.self.in, pre ? &w.self.norm : nullptr, x_.ptr, rows, nullptr, -1, qkv_.ptr, xq);
    launch_attention_encoder(qkv_.ptr, lens_d, batch, static_cast<int>(S), mc_.num_heads, mc_.head_dim, scale, ctx_.ptr, dtype_,
                             stream_);
    dense(w.self.out, nullptr, ctx_.ptr, rows, x_.ptr, -1, x_.ptr);
    xq = !pre && post_norm(w.self.norm, x_.ptr, rows, &w.ffn.ff1);
    dense(w.ffn.ff1, pre ? &w.ffn.norm : nullptr, x_.ptr, rows, nullptr, mc_.enc_activation, h_.ptr, xq);
    dense(w.ffn.ff2, nullptr, h_.ptr, rows, x_.ptr, -1, x_.ptr);
    xq = !pre && post_norm(w.ffn.norm, x_.ptr, rows, l + 1 < mc_.enc_layers ? &enc_[l + 1].self.in : nullptr);
  }
  if (mc_.has_enc_final_norm)
    launch_layer_norm(x_.ptr, enc_norm_.gamma.ptr, enc_norm_.beta.ptr, rows, d, mc_.eps, memory_.ptr, nullptr, nullptr, true, dtype_,
                      stream_);
  else
    CT2_CUDA_CHECK(cudaMemcpyAsync(memory_.ptr, x_.ptr, rows * d * dtype_size(dtype_), cudaMemcpyDeviceToDevice, stream_));
}

// the memory keys / values of every decoder layer, once per batch (cached_attn_keys / values, attention.cc:385-428)
void Translator::project_memory(int64_t batch, int64_t S) {
  for (int l = 0; l < mc_.dec_layers; ++l)
    dense(dec_[l].cross.kv, nullptr, memory_.ptr, batch * S, nullptr, -1, mem_kv_[l].ptr);
}

// TransformerDecoder::decode for one target position of every beam row (transformer.cc:621-871)
void Translator::decoder_step(int64_t rows, int beam, int64_t batch, int64_t S) {
  (void)batch;
  const int64_t d = mc_.d_model;
  const float scale = 1.f / std::sqrt(static_cast<float>(mc_.head_dim));
  const bool pre = mc_.dec_pre_norm;
  const int32_t* step_ptr = beam_.counters.as<int32_t>();
  launch_embed_pos(dec_emb_.weight.ptr, dec_emb_.kind == DenseWeights::INT8 ? dec_emb_.scale.as<float>() : nullptr,
                   beam_.next_ids.as<int32_t>(), rows, d, mc_.dec_emb_scale, dec_pos_.ptr, 1, step_ptr, mc_.start_from_zero_embedding, x_.ptr, dtype_,
                   stream_);
  bool xq = false;                                 // xq_ / xs_ already hold Quantize(x_) (left by a post-norm launch)
  for (int l = 0; l < mc_.dec_layers; ++l) {
    DecoderLayerWeights& w = dec_[l];
    dense(w.self.in, pre ? &w.self.norm : nullptr, x_.ptr, rows, nullptr, -1, qkv_.ptr, xq);
    launch_attention_beam_self(qkv_.ptr, self_k_[l].ptr, self_v_[l].ptr, beam_.anc.as<int32_t>(), step_ptr, rows,
                               static_cast<int>(cap_steps_), mc_.num_heads, mc_.head_dim, scale, ctx_.ptr, dtype_, stream_);
    dense(w.self.out, nullptr, ctx_.ptr, rows, x_.ptr, -1, x_.ptr);
    xq = !pre && post_norm(w.self.norm, x_.ptr, rows, &w.cross.in);
    dense(w.cross.in, pre ? &w.cross.norm : nullptr, x_.ptr, rows, nullptr, -1, q_.ptr, xq);
    launch_attention_cross(q_.ptr, mem_kv_[l].ptr, src_lens_.as<int32_t>(), rows, beam, static_cast<int>(S), mc_.num_heads,
                           mc_.head_dim, scale, ctx_.ptr, dtype_, stream_);
    dense(w.cross.out, nullptr, ctx_.ptr, rows, x_.ptr, -1, x_.ptr);
    xq = !pre && post_norm(w.cross.norm, x_.ptr, rows, &w.ffn.ff1);
    dense(w.ffn.ff1, pre ? &w.ffn.norm : nullptr, x_.ptr, rows, nullptr, mc_.dec_activation, h_.ptr, xq);
    dense(w.ffn.ff2, nullptr, h_.ptr, rows, x_.ptr, -1, x_.ptr);
    xq = !pre && post_norm(w.ffn.norm, x_.ptr, rows,
                           l + 1 < mc_.dec_layers ? &dec_[l + 1].self.in : (mc_.has_dec_final_norm ? nullptr : &projection_));
  }
  dense(projection_, mc_.has_dec_final_norm ? &dec_norm_ : nullptr, x_.ptr, rows, nullptr, -1, logits_.ptr, xq, logits_ld_);
}

void Translator::launch_or_capture_step(const BeamState& bs, int64_t S) {
  const int64_t rows = static_cast<int64_t>(bs.batch) * bs.beam;
  if (!use_graph_) {
    decoder_step(rows, bs.beam, bs.batch, S);
    beam_.step(logits_.ptr, bs, dtype_, stream_);
    return;
  }
  if (!graph_) {
    cudaGraph_t g = nullptr;
    const int64_t before = g_kernel_launches.load();
    CT2_CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    try {
      decoder_step(rows, bs.beam, bs.batch, S);
      beam_.step(logits_.ptr, bs, dtype_, stream_);
    } catch (...) {
      cudaStreamEndCapture(stream_, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CT2_CUDA_CHECK(cudaStreamEndCapture(stream_, &g));
    g_kernel_launches.store(before);
    size_t n = 0;
    cudaGraphGetNodes(g, nullptr, &n);
    graph_nodes_ = static_cast<int64_t>(n);
    CT2_CUDA_CHECK(cudaGraphInstantiate(&graph_, g, 0));
    cudaGraphDestroy(g);
  }
  CT2_CUDA_CHECK(cudaGraphLaunch(graph_, stream_));
  count_launch(static_cast<int>(graph_nodes_));
}

// =============================================================================================
// the search (shared by translate_batch and Whisper::generate)
// =============================================================================================
// the decoding loop: one captured step per position; the host only polls the "finished entries" counter
void Translator::run_search(const BeamState& bs, int64_t S, int64_t first_check) {
  // everything the captured step bakes in (kernel arguments are values)
  std::vector<int64_t> key = {bs.batch, bs.beam, S, bs.vocab_ld, bs.stride, bs.max_steps, bs.min_length, bs.max_hyp, bs.max_candidates,
                              bs.num_hypotheses, bs.early_exit, bs.num_end, bs.start_step, bs.include_eos, bs.num_disable,
                              bs.num_begin, bs.ts_begin, bs.ts_end, bs.ts_eot, bs.ts_no_timestamps, bs.ts_max_initial};
  if (key != graph_key_) {
    if (graph_) {
      cudaGraphExecDestroy(graph_);
      graph_ = nullptr;
    }
    graph_key_ = key;
  }
  const char* poll_env = std::getenv("CT2B200_EOS_POLL");
  const int64_t poll = std::max<int64_t>(1, poll_env ? std::atoll(poll_env) : 4);
  int32_t* hfin = beam_.host;
  for (int64_t s = 0; s < bs.max_steps; ++s) {
    launch_or_capture_step(bs, S);
    if (s + 1 == bs.max_steps) break;
    if (s >= first_check && (s - first_check) % poll == poll - 1) {
      CT2_CUDA_CHECK(cudaMemcpyAsync(hfin, bs.num_finished, 4, cudaMemcpyDeviceToHost, stream_));
      CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
      if (*hfin >= bs.batch) break;
    }
  }
}

// =============================================================================================
// Translator::translate_batch
// =============================================================================================
std::vector<TranslationHypotheses> Translator::translate(const TranslationRequest& r) {
  std::lock_guard<std::mutex> lock(mu_);
  CT2_REQUIRE(!mc_.whisper, "this is a Whisper model: use whisper_generate");
  const int64_t B = r.batch, S = r.max_source_len, L = r.max_decoding_length;
  const int beam = r.beam_size;
  CT2_REQUIRE(B > 0 && S > 0, "translate_batch: empty batch");
  CT2_REQUIRE(beam >= 1 && beam <= 32, "beam_size must be in [1, 32]");
  CT2_REQUIRE(r.num_hypotheses >= 1 && r.num_hypotheses <= beam, "num_hypotheses must be in [1, beam_size]");   // decoding.cc:1046-1048
  CT2_REQUIRE(L >= 1 && r.min_decoding_length <= L, "min_decoding_length is greater than max_decoding_length");
  CT2_REQUIRE(r.patience > 0.f && r.patience <= 2.f, "patience must be in (0, 2]");
  CT2_REQUIRE(r.end_ids.size() <= 64, "at most 64 end tokens");
  CT2_REQUIRE(static_cast<int64_t>(2) * beam <= static_cast<int64_t>(beam) * mc_.tgt_vocab, "beam_size exceeds the vocabulary");
  for (int64_t b = 0; b < B; ++b) {
    CT2_REQUIRE(r.source_lens[b] >= 1 && r.source_lens[b] <= S, "translate_batch: source lengths must be in [1, max_source_len]");
    for (int64_t t = 0; t < r.source_lens[b]; ++t) {
      const int32_t id = r.source_ids[b * S + t];
      CT2_REQUIRE(id >= 0 && id < mc_.src_vocab, "translate_batch: source id out of range");
    }
  }
  ensure_arena(B, S, beam, L);

  // ---- inputs ----
  int32_t* hp = host_pinned_;
  for (int64_t b = 0; b < B; ++b)
    for (int64_t t = 0; t < S; ++t) hp[b * S + t] = t < r.source_lens[b] ? r.source_ids[b * S + t] : 0;
  int32_t* hl = hp + B * S;
  for (int64_t b = 0; b < B; ++b) hl[b] = r.source_lens[b];
  int32_t* hend = hl + B;
  for (size_t i = 0; i < r.end_ids.size(); ++i) hend[i] = r.end_ids[i];
  CT2_CUDA_CHECK(cudaMemcpyAsync(src_ids_.ptr, hp, B * S * 4, cudaMemcpyHostToDevice, stream_));
  CT2_CUDA_CHECK(cudaMemcpyAsync(src_lens_.ptr, hl, B * 4, cudaMemcpyHostToDevice, stream_));
  if (!r.end_ids.empty())
    CT2_CUDA_CHECK(cudaMemcpyAsync(beam_.end_ids.ptr, hend, r.end_ids.size() * 4, cudaMemcpyHostToDevice, stream_));

  // ---- encoder + memory projections ----
  run_encoder(B, S);
  project_memory(B, S);

  // ---- beam search: the earliest step at which an entry can be complete is min_decoding_length ----
  BeamState bs = beam_.state(B, beam, mc_.tgt_vocab, L, r.min_decoding_length, r.patience, r.length_penalty, r.num_hypotheses,
                             static_cast<int>(r.end_ids.size()));
  set_logits_ld(bs);
  beam_.reset(bs, r.start_id, dtype_, stream_);
  run_search(bs, S, std::max<int64_t>(0, r.min_decoding_length));
  return beam_.collect(bs, r.length_penalty, r.num_hypotheses, r.return_end_token ? std::vector<int32_t>{} : r.end_ids, stream_);
}

void Translator::encode(const int32_t* ids_h, const int32_t* lens_h, int64_t batch, int64_t S, float* memory_h) {
  std::lock_guard<std::mutex> lock(mu_);
  CT2_REQUIRE(!mc_.whisper, "this is a Whisper model: use whisper_encode");
  CT2_REQUIRE(batch > 0 && S > 0, "encode: empty batch");
  ensure_arena(batch, S, 1, 1);
  int32_t* hp = host_pinned_;
  for (int64_t b = 0; b < batch; ++b) {
    CT2_REQUIRE(lens_h[b] >= 1 && lens_h[b] <= S, "encode: source lengths must be in [1, max_source_len]");
    for (int64_t t = 0; t < S; ++t) hp[b * S + t] = t < lens_h[b] ? ids_h[b * S + t] : 0;
  }
  CT2_CUDA_CHECK(cudaMemcpyAsync(src_ids_.ptr, hp, batch * S * 4, cudaMemcpyHostToDevice, stream_));
  CT2_CUDA_CHECK(cudaMemcpyAsync(src_lens_.ptr, lens_h, batch * 4, cudaMemcpyHostToDevice, stream_));
  run_encoder(batch, S);
  DeviceBuffer f32(static_cast<size_t>(batch) * S * mc_.d_model * 4);
  launch_convert_to_f32(memory_.ptr, batch * S * mc_.d_model, f32.as<float>(), dtype_, stream_);
  CT2_CUDA_CHECK(cudaMemcpyAsync(memory_h, f32.ptr, f32.bytes, cudaMemcpyDeviceToHost, stream_));
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Translator::bench(int64_t batch, int64_t source_len, int beam, int64_t steps, int64_t warmup, float* encode_ms,
                       float* decode_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lock(mu_);
  CT2_REQUIRE(!mc_.whisper, "bench_translate serves Translator models");
  const int64_t L = steps + warmup;
  ensure_arena(batch, source_len, beam, L);
  std::vector<int32_t> ids(batch * source_len), lens(batch, static_cast<int32_t>(source_len));
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = static_cast<int32_t>((7919ull * i + 3) % mc_.src_vocab);
  CT2_CUDA_CHECK(cudaMemcpy(src_ids_.ptr, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
  CT2_CUDA_CHECK(cudaMemcpy(src_lens_.ptr, lens.data(), lens.size() * 4, cudaMemcpyHostToDevice));
  CT2_CUDA_CHECK(cudaDeviceSynchronize());      // pageable H2D: the DMA may still be running when cudaMemcpy returns (see upload())
  BeamState bs = beam_.state(batch, beam, mc_.tgt_vocab, L, 0, 1.f, 1.f, 1, 0);   // no end token: nothing finishes early
  set_logits_ld(bs);
  std::vector<int64_t> key = {-1, batch, beam, source_len, bs.stride, L};
  if (key != graph_key_) {
    if (graph_) {
      cudaGraphExecDestroy(graph_);
      graph_ = nullptr;
    }
    graph_key_ = key;
  }
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventCreate(&e2);
  cudaEventCreate(&e3);
  run_encoder(batch, source_len);      // warm-up (first-use kernel configuration)
  project_memory(batch, source_len);
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaEventRecord(e0, stream_);
  run_encoder(batch, source_len);
  project_memory(batch, source_len);
  cudaEventRecord(e1, stream_);
  beam_.reset(bs, 1, dtype_, stream_);
  for (int64_t s = 0; s < warmup; ++s) launch_or_capture_step(bs, source_len);
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
  const int64_t l0 = g_kernel_launches.load();
  cudaProfilerStart();
  cudaEventRecord(e2, stream_);
  for (int64_t s = 0; s < steps; ++s) launch_or_capture_step(bs, source_len);
  cudaEventRecord(e3, stream_);
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaProfilerStop();
  *launches = g_kernel_launches.load() - l0;
  cudaEventElapsedTime(encode_ms, e0, e1);
  cudaEventElapsedTime(decode_ms, e2, e3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  cudaEventDestroy(e3);
}

// =============================================================================================
// Whisper (src/models/whisper.cc, src/layers/whisper.cc)
// =============================================================================================
// WhisperEncoder::operator(): conv1 + GELU, conv2 (stride 2) + GELU as im2col + float Dense (the GEMM output is already the
// transposed [batch, frames / 2, d] layout), stored positions, pre-norm GELU layers, LayerNorm
void Translator::run_whisper_encoder(int64_t batch, int64_t frames) {
  const int64_t d = mc_.d_model, S = (frames + 2 - 3) / 2 + 1;
  launch_im2col(features_.ptr, true, batch, mc_.n_mels, frames, frames, 3, 1, 1, true, cols_.ptr, dtype_, stream_);
  gemm_float(cols_.ptr, conv1_.weight.ptr, conv1_.bias.ptr, nullptr, CT2B200_ACT_GELU, batch * frames, d, mc_.n_mels * 3,
             conv_out_.ptr, dtype_, stream_);
  launch_im2col(conv_out_.ptr, false, batch, d, frames, S, 3, 2, 1, false, cols_.ptr, dtype_, stream_);
  gemm_float(cols_.ptr, conv2_.weight.ptr, conv2_.bias.ptr, nullptr, CT2B200_ACT_GELU, batch * S, d, d * 3, x_.ptr, dtype_, stream_);
  launch_add_positions(x_.ptr, enc_pos_.ptr, batch * S, S, d, dtype_, stream_);
  run_encoder_layers(batch, S, nullptr);
}

void Translator::whisper_encode(const float* features_h, int64_t batch, int64_t frames, float* memory_h) {
  std::lock_guard<std::mutex> lock(mu_);
  CT2_REQUIRE(mc_.whisper, "whisper_encode needs a Whisper model");
  const int64_t S = (frames + 2 - 3) / 2 + 1;
  CT2_REQUIRE(batch > 0 && frames >= 2 && S <= mc_.max_frames, "Invalid input features shape: too many frames for the encoder");
  ensure_arena(batch, S, 1, 1);
  CT2_CUDA_CHECK(cudaMemcpyAsync(features_.ptr, features_h, batch * mc_.n_mels * frames * 4, cudaMemcpyHostToDevice, stream_));
  run_whisper_encoder(batch, frames);
  DeviceBuffer f32(static_cast<size_t>(batch) * S * mc_.d_model * 4);
  launch_convert_to_f32(memory_.ptr, batch * S * mc_.d_model, f32.as<float>(), dtype_, stream_);
  CT2_CUDA_CHECK(cudaMemcpyAsync(memory_h, f32.ptr, f32.bytes, cudaMemcpyDeviceToHost, stream_));
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
}

std::vector<TranslationHypotheses> Translator::whisper_generate(const WhisperRequest& r, float* no_speech_h) {
  std::lock_guard<std::mutex> lock(mu_);
  CT2_REQUIRE(mc_.whisper, "whisper_generate needs a Whisper model");
  const int64_t B = r.batch, P = r.prompt_len, frames = r.frames;
  const int64_t S = (frames + 2 - 3) / 2 + 1;
  const int beam = r.beam_size;
  CT2_REQUIRE(B > 0 && frames >= 2 && S <= mc_.max_frames, "Invalid input features shape: too many frames for the encoder");
  CT2_REQUIRE(beam >= 1 && beam <= 32, "beam_size must be in [1, 32]");
  CT2_REQUIRE(r.num_hypotheses >= 1 && r.num_hypotheses <= beam, "num_hypotheses must be in [1, beam_size]");
  CT2_REQUIRE(r.patience > 0.f && r.patience <= 2.f, "patience must be in (0, 2]");
  CT2_REQUIRE(r.suppress_ids.size() + r.suppress_ids_begin.size() <= 4096, "too many suppressed tokens");
  // check_prompts (whisper.cc:168-197): <|startoftranscript|> at the same position in every prompt and the same number of
  // task tokens after it; this engine also requires the prompt to END with the task tokens (no text after them)
  CT2_REQUIRE(P >= 1 && P <= 16, "the prompt must hold <|startoftranscript|> and the task tokens (1 to 16 tokens)");
  CT2_REQUIRE(r.no_timestamps_id > r.sot_id && r.no_timestamps_id < mc_.tgt_vocab - 1, "no_timestamps_id must follow sot_id");
  int64_t sot_index = -1;
  for (int64_t b = 0; b < B; ++b) {
    int64_t idx = -1;
    for (int64_t t = 0; t < P; ++t) {
      const int32_t id = r.prompts[b * P + t];
      CT2_REQUIRE(id >= 0 && id < mc_.tgt_vocab, "prompt id out of range");
      if (id == r.sot_id && idx < 0) idx = t;
    }
    CT2_REQUIRE(idx >= 0, "<|startoftranscript|> token was not found in the prompt");
    CT2_REQUIRE(sot_index < 0 || idx == sot_index, "<|startoftranscript|> must be at the same position in all prompts");
    sot_index = idx;
    for (int64_t t = idx; t < P; ++t)                  // get_prompt_length (whisper.cc:156-166)
      CT2_REQUIRE(r.prompts[b * P + t] >= r.sot_id && r.prompts[b * P + t] <= r.no_timestamps_id,
                  "text after the task tokens (a decoding prefix) is not supported");
  }
  bool timestamps = r.prompts[P - 1] != r.no_timestamps_id;             // whisper.cc:325, decided on the first prompt
  const int64_t start_step = P - 1;
  const int64_t steps = std::min<int64_t>(r.max_length / 2, r.max_length - start_step);      // whisper.cc:299
  CT2_REQUIRE(steps >= 1, "max_length is too small for the prompt");
  ensure_arena(B, S, beam, start_step + steps);
  const int64_t N = B * beam;

  // ---- inputs ----
  CT2_CUDA_CHECK(cudaMemcpyAsync(features_.ptr, r.features, B * mc_.n_mels * frames * 4, cudaMemcpyHostToDevice, stream_));
  int32_t* hp = host_pinned_;
  for (int64_t t = 0; t < P; ++t)                      // forced inputs [P, N]: prompt token t of the row's batch entry
    for (int64_t n = 0; n < N; ++n) hp[t * N + n] = r.prompts[(n / beam) * P + t];
  int32_t* hs = hp + P * N;
  for (size_t i = 0; i < r.suppress_ids.size(); ++i) hs[i] = r.suppress_ids[i];
  for (size_t i = 0; i < r.suppress_ids_begin.size(); ++i) hs[r.suppress_ids.size() + i] = r.suppress_ids_begin[i];
  hs[r.suppress_ids.size() + r.suppress_ids_begin.size()] = r.eot_id;
  const size_t nsup = r.suppress_ids.size() + r.suppress_ids_begin.size();
  forced_d_.alloc(P * N * 4);
  suppress_d_.alloc((nsup + 1) * 4);
  CT2_CUDA_CHECK(cudaMemcpyAsync(forced_d_.ptr, hp, P * N * 4, cudaMemcpyHostToDevice, stream_));
  CT2_CUDA_CHECK(cudaMemcpyAsync(suppress_d_.ptr, hs, (nsup + 1) * 4, cudaMemcpyHostToDevice, stream_));
  CT2_CUDA_CHECK(cudaMemcpyAsync(beam_.end_ids.ptr, suppress_d_.as<int32_t>() + nsup, 4, cudaMemcpyDeviceToDevice, stream_));
  std::vector<int32_t> lens(B, static_cast<int32_t>(S));
  CT2_CUDA_CHECK(cudaMemcpyAsync(src_lens_.ptr, lens.data(), B * 4, cudaMemcpyHostToDevice, stream_));
  CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));      // `lens` and the pinned staging are reused below

  // ---- encoder + memory projections ----
  run_whisper_encoder(B, frames);
  project_memory(B, S);

  // ---- prompt: WhisperDecoder::forward_prompt on prompt[:-1], one position per step, no search ----
  BeamState bs = beam_.state(B, beam, mc_.tgt_vocab, steps, 0, r.patience, r.length_penalty, r.num_hypotheses, 1);
  set_logits_ld(bs);
  bs.start_step = static_cast<int>(start_step);
  bs.include_eos = 0;                                  // whisper.cc:309
  bs.num_disable = static_cast<int>(r.suppress_ids.size());
  bs.num_begin = static_cast<int>(r.suppress_ids_begin.size());
  bs.disable_ids = suppress_d_.as<int32_t>();
  bs.disable_begin = suppress_d_.as<int32_t>() + r.suppress_ids.size();
  if (timestamps) {
    bs.ts_begin = r.no_timestamps_id + 1;
    bs.ts_end = static_cast<int>(mc_.tgt_vocab) - 1;
    bs.ts_eot = r.eot_id;
    bs.ts_no_timestamps = r.no_timestamps_id;
    bs.ts_max_initial = bs.ts_begin + r.max_initial_timestamp_index;
  }
  beam_.reset(bs, r.prompts[0], dtype_, stream_);
  CT2_CUDA_CHECK(cudaMemcpyAsync(beam_.next_ids.ptr, forced_d_.ptr, N * 4, cudaMemcpyDeviceToDevice, stream_));
  no_speech_d_.alloc(B * 4);
  for (int64_t t = 0; t < start_step; ++t) {
    decoder_step(N, beam, B, S);
    if (no_speech_h && t == sot_index) {
      CT2_REQUIRE(r.no_speech_id >= 0, "return_no_speech_prob needs the id of <|nospeech|>");
      launch_token_prob(logits_.ptr, B, mc_.tgt_vocab, static_cast<int64_t>(beam) * logits_ld_, r.no_speech_id,
                        no_speech_d_.as<float>(), dtype_, stream_);
    }
    launch_beam_force(bs, forced_d_.as<int32_t>() + (t + 1) * N, stream_);
  }
  CT2_REQUIRE(!no_speech_h || sot_index < start_step, "return_no_speech_prob with <|startoftranscript|> as the last prompt "
                                                      "token is not supported");

  // ---- search ----
  run_search(bs, S, 0);
  if (no_speech_h) {
    CT2_CUDA_CHECK(cudaMemcpyAsync(no_speech_h, no_speech_d_.ptr, B * 4, cudaMemcpyDeviceToHost, stream_));
    CT2_CUDA_CHECK(cudaStreamSynchronize(stream_));
  }
  return beam_.collect(bs, r.length_penalty, r.num_hypotheses, {}, stream_);
}

}  // namespace ct2b200

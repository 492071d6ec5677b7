// translator.h — the encoder-decoder path behind ctranslate2::Translator (BASELINE config 2): model loading for
// TransformerSpec directories (src/models/transformer.cc), TransformerEncoder (src/layers/transformer.cc:405-471),
// TransformerDecoder with cross-attention (:621-871, attention.cc:371-440) and BeamSearch::search (src/decoding.cc:425-720)
// resident on the device: the whole decoding step — embeddings to beam bookkeeping — is one CUDA graph, beams are
// reordered by an index remap of the K/V arena, and the host only polls a "finished entries" counter.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "beam.h"
#include "engine.h"

namespace ct2b200 {

struct Seq2SeqConfig {
  int enc_layers = 0, dec_layers = 0, num_heads = 8, head_dim = 0;
  int64_t d_model = 0, ffn_dim = 0, src_vocab = 0, tgt_vocab = 0;
  bool enc_pre_norm = true, dec_pre_norm = true;
  int enc_activation = CT2B200_ACT_RELU, dec_activation = CT2B200_ACT_RELU;
  float enc_emb_scale = 0.f, dec_emb_scale = 0.f;   // 0 = embeddings are not scaled
  float eps = 1e-5f;
  bool round_before_cast = true;                     // binary version >= 5 (model.h:87-89)
  bool has_enc_final_norm = false, has_dec_final_norm = false;
  bool start_from_zero_embedding = false;            // Marian / OPUS-MT decoders (transformer.cc:637-640)
  bool whisper = false;                              // WhisperSpec: Conv1D front-end instead of source embeddings
  int64_t n_mels = 0, max_frames = 0;                // Whisper: input channels, encoder positions (frames / 2)
  int64_t weight_bytes = 0;
  std::string weights;                               // storage type of the linear layers
};

Seq2SeqConfig parse_seq2seq_config(const ModelFile& file);   // host only

struct NormWeights {
  DeviceBuffer gamma, beta;
};
struct AttentionWeights {
  NormWeights norm;
  DenseWeights in;      // self-attention: fused q|k|v; cross-attention: q
  DenseWeights kv;      // cross-attention: fused k|v of the memory
  DenseWeights out;
};
struct FfnWeights {
  NormWeights norm;
  DenseWeights ff1, ff2;
};
struct EncoderLayerWeights {
  AttentionWeights self;
  FfnWeights ffn;
};
struct DecoderLayerWeights {
  AttentionWeights self, cross;
  FfnWeights ffn;
};

struct TranslationRequest {
  const int32_t* source_ids = nullptr;    // host [batch, max_source_len], right-padded
  const int32_t* source_lens = nullptr;   // host [batch]
  int64_t batch = 0, max_source_len = 0;
  int beam_size = 2;                      // TranslationOptions defaults (include/ctranslate2/translation.h)
  float patience = 1.f;
  float length_penalty = 1.f;
  int64_t max_decoding_length = 256, min_decoding_length = 1;
  int num_hypotheses = 1;
  int32_t start_id = 1;                   // decoder start token (<s>)
  std::vector<int32_t> end_ids;           // normally {</s>}
  bool return_end_token = false;
};

// models::Whisper::generate (include/ctranslate2/models/whisper.h:11-60, src/models/whisper.cc:232-390), prompts made of
// previous-text tokens, <|startoftranscript|> and the task tokens (no text after them); the timestamp rules
// (whisper.cc:742-860) apply unless the last task token is <|notimestamps|>
struct WhisperRequest {
  const float* features = nullptr;        // host [batch, n_mels, frames] f32
  int64_t batch = 0, frames = 0;
  const int32_t* prompts = nullptr;       // host [batch, prompt_len]
  int64_t prompt_len = 0;
  int beam_size = 5;
  float patience = 1.f, length_penalty = 1.f;
  int64_t max_length = 448;
  int num_hypotheses = 1;
  std::vector<int32_t> suppress_ids, suppress_ids_begin;
  int32_t sot_id = 0, eot_id = 0, no_speech_id = -1, no_timestamps_id = -1;
  int max_initial_timestamp_index = 50;
  bool return_no_speech_prob = false;
};


class Translator {
 public:
  Translator(const std::string& model_dir, const ct2b200_generator_config& cfg);
  ~Translator();
  const Seq2SeqConfig& config() const { return mc_; }
  int dtype() const { return dtype_; }

  // Translator::translate_batch on token ids
  std::vector<TranslationHypotheses> translate(const TranslationRequest& req);
  // TransformerEncoder::operator(): memory_h [batch, max_source_len, d_model] f32 host
  void encode(const int32_t* ids_h, const int32_t* lens_h, int64_t batch, int64_t max_source_len, float* memory_h);
  // WhisperEncoder::operator(): features_h [batch, n_mels, frames] f32 -> memory_h [batch, frames / 2, d_model] f32
  void whisper_encode(const float* features_h, int64_t batch, int64_t frames, float* memory_h);
  // models::Whisper::generate; no_speech_h [batch] or null
  std::vector<TranslationHypotheses> whisper_generate(const WhisperRequest& req, float* no_speech_h);
  // device-timed phases for bench.py: encoder pass, then `steps` decoding steps of batch * beam rows
  void bench(int64_t batch, int64_t source_len, int beam, int64_t steps, int64_t warmup, float* encode_ms, float* decode_ms,
             int64_t* launches);

 private:
  void load_dense(const ModelFile& f, const std::string& prefix, DenseWeights& w);
  void load_norm(const ModelFile& f, const std::string& prefix, NormWeights& n);
  void ensure_arena(int64_t batch, int64_t src_len, int beam, int64_t max_steps);
  // Dense on T rows (quantizes them for int8 weights); `pre` = the LayerNorm applied first (fused with the quantization)
  void dense(const DenseWeights& w, const NormWeights* pre, const void* x, int64_t rows, const void* residual, int act, void* y,
             bool prequantized = false, int64_t ldy = 0);
  void set_logits_ld(BeamState& bs);
  bool post_norm(const NormWeights& n, void* x, int64_t rows, const DenseWeights* next);
  void run_encoder(int64_t batch, int64_t S);
  void run_encoder_layers(int64_t batch, int64_t S, const int32_t* lens_d);
  void run_whisper_encoder(int64_t batch, int64_t frames);
  void run_search(const BeamState& bs, int64_t S, int64_t first_check);
  void project_memory(int64_t batch, int64_t S);
  void decoder_step(int64_t rows, int beam, int64_t batch, int64_t S);
  void launch_or_capture_step(const BeamState& bs, int64_t S);

  std::mutex mu_;                // translate / encode / bench are serialised per translator
  Seq2SeqConfig mc_;
  int dtype_ = CT2B200_F32, device_ = 0, weight_type_ = CT2B200_WEIGHTS_STORED, sm_count_ = 148;
  bool use_graph_ = true;
  cudaStream_t stream_ = nullptr;

  DenseWeights enc_emb_, dec_emb_, projection_;
  DenseWeights conv1_, conv2_;   // Whisper: [d, n_mels * 3] / [d, d * 3] in T (+ bias)
  DeviceBuffer enc_pos_, dec_pos_;
  int64_t enc_positions_ = 0, dec_positions_ = 0;
  NormWeights enc_norm_, dec_norm_;
  std::vector<EncoderLayerWeights> enc_;
  std::vector<DecoderLayerWeights> dec_;

  // arena (grown on demand)
  int64_t cap_batch_ = 0, cap_src_ = 0, cap_rows_ = 0, cap_steps_ = 0;
  int cap_beam_ = 0;
  DeviceBuffer src_ids_, src_lens_, x_, xn_, xq_, xs_, qkv_, ctx_, h_, q_, memory_;
  std::vector<DeviceBuffer> mem_kv_, self_k_, self_v_;
  DeviceBuffer logits_;
  int64_t logits_ld_ = 0;                 // row stride of logits_ for the current search (set_logits_ld)
  BeamSearchArena beam_;         // search state: next ids, scores, histories, ancestry, hypotheses, counters
  DeviceBuffer features_, cols_, conv_out_, suppress_d_, forced_d_, no_speech_d_;   // Whisper
  int64_t cap_frames_ = 0;
  int32_t* host_pinned_ = nullptr;
  size_t host_pinned_elems_ = 0;

  cudaGraphExec_t graph_ = nullptr;
  int64_t graph_nodes_ = 0;
  std::vector<int64_t> graph_key_;
};

}  // namespace ct2b200

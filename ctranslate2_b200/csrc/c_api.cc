// c_api.cc — the extern "C" boundary declared in include/ct2b200.h.  Every entry point converts C++
// exceptions into an error code + thread-local message (the reference surfaces std::invalid_argument /
// std::runtime_error through std::future::get(), src/cuda/utils.h:51-96).
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "common.cuh"
#include "host/beam.h"
#include "host/engine.h"
#include "host/translator.h"
#include "kernels/beam_decide.h"
#include "kernels/kernels.h"

using namespace ct2b200;

namespace {
thread_local std::string g_error;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const InvalidArgument& e) {
    g_error = std::string("invalid argument: ") + e.what();
    return 2;
  } catch (const std::invalid_argument& e) {
    g_error = std::string("invalid argument: ") + e.what();
    return 2;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

void require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    throw std::runtime_error("no CUDA device: ct2b200 has no CPU fallback");
}
}  // namespace

struct ct2b200_generator {
  std::unique_ptr<Generator> impl;
};
struct ct2b200_translator {
  std::unique_ptr<Translator> impl;
};

extern "C" {

CT2B200_API const char* ct2b200_last_error(void) { return g_error.c_str(); }
CT2B200_API const char* ct2b200_version(void) { return "0.1.0 (sm_100a)"; }
CT2B200_API int64_t ct2b200_kernel_launch_count(void) { return g_kernel_launches.load(); }

CT2B200_API int ct2b200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  return guarded([&] {
    require_device();
    cudaDeviceProp p;
    CT2_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
  });
}

CT2B200_API int ct2b200_quantize_rows(const void* x, int dtype, int64_t rows, int64_t cols, int round_before_cast, int8_t* q,
                          float* scale, void* stream) {
  return guarded([&] {
    require_device();
    launch_quantize_rows(x, dtype, rows, cols, round_before_cast != 0, q, scale, S(stream));
  });
}

CT2B200_API int ct2b200_gemm_s8(const int8_t* a, const int8_t* b, int64_t m, int64_t n, int64_t k, int32_t* c, int impl,
                    void* stream) {
  return guarded([&] {
    require_device();
    DenseEpilogue e{nullptr, nullptr, nullptr, nullptr, nullptr, c, -1, n};
    gemm_s8(a, b, m, n, k, e, CT2B200_F32, impl, S(stream));
  });
}

CT2B200_API int ct2b200_dequantize_gemm_output(const int32_t* c, const float* a_scale, const float* b_scale, const void* bias,
                                   int act, int64_t m, int64_t n, void* y, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    DenseEpilogue e{a_scale, b_scale, bias, nullptr, y, nullptr, act, n};
    launch_dequantize_gemm_output(c, e, m, n, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_dequantize_rows(const int8_t* x, const float* scale, int64_t rows, int64_t cols, void* y, int dtype,
                            void* stream) {
  return guarded([&] {
    require_device();
    launch_dequantize_rows(x, scale, rows, cols, y, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_dense_s8(const int8_t* xq, const float* x_scale, const int8_t* w, const float* w_scale, const void* bias,
                     const void* residual, int act, int64_t m, int64_t n, int64_t k, void* y, int dtype, int impl,
                     void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(x_scale && w_scale, "dense_s8: scales are required");
    DenseEpilogue e{x_scale, w_scale, bias, residual, y, nullptr, act, n};
    gemm_s8(xq, w, m, n, k, e, dtype, impl, S(stream));
  });
}

CT2B200_API int ct2b200_dense_s8_glu(const int8_t* xq, const float* x_scale, const int8_t* w_gate, const float* w_gate_scale,
                         const int8_t* w_up, const float* w_up_scale, int act, int64_t m, int64_t n, int64_t k, void* h,
                         int dtype, int impl, void* stream) {
  return guarded([&] {
    require_device();
    GluEpilogue g{x_scale, w_gate_scale, w_up_scale, h, act, n};
    gemm_s8_glu(xq, w_gate, w_up, m, n, k, g, dtype, impl, S(stream));
  });
}

namespace {
void rows_to_int8(const void* x, const void* gamma, float eps, int64_t m, int64_t k, int dtype, int8_t* xq, float* xs,
                  cudaStream_t st) {
  if (gamma) launch_rms_norm(gamma, x, m, k, eps, false, nullptr, xq, xs, dtype, st);
  else launch_quantize_rows(x, dtype, m, k, true, xq, xs, st);
}
}  // namespace

CT2B200_API int ct2b200_dense_s8_rows(const void* x, const void* gamma, float eps, const int8_t* w, const float* w_scale,
                          const void* bias, const void* residual, int act, int64_t m, int64_t n, int64_t k, void* y,
                          int dtype, int8_t* xq, float* x_scale, unsigned* barrier, void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(x && w && w_scale && xq && x_scale && barrier, "dense_s8_rows: null argument");
    DenseEpilogue e{x_scale, w_scale, bias, residual, y, nullptr, act, n};
    RowPre pre{gamma ? 2 : 1, x, gamma, eps, barrier};
    if (m <= 64 && gemm_s8_decode(xq, w, m, n, k, e, dtype, S(stream), &pre)) return;
    rows_to_int8(x, gamma, eps, m, k, dtype, xq, x_scale, S(stream));
    gemm_s8(xq, w, m, n, k, e, dtype, CT2B200_GEMM_AUTO, S(stream));
  });
}

CT2B200_API int ct2b200_dense_s8_glu_rows(const void* x, const void* gamma, float eps, const int8_t* w_gate,
                              const float* w_gate_scale, const int8_t* w_up, const float* w_up_scale, int act, int64_t m,
                              int64_t n, int64_t k, void* h, int dtype, int8_t* xq, float* x_scale, unsigned* barrier,
                              void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(x && w_gate && w_up && xq && x_scale && barrier, "dense_s8_glu_rows: null argument");
    GluEpilogue g{x_scale, w_gate_scale, w_up_scale, h, act, n};
    RowPre pre{gamma ? 2 : 1, x, gamma, eps, barrier};
    if (m <= 64 && gemm_s8_glu_decode(xq, w_gate, w_up, m, n, k, g, dtype, S(stream), &pre)) return;
    rows_to_int8(x, gamma, eps, m, k, dtype, xq, x_scale, S(stream));
    gemm_s8_glu(xq, w_gate, w_up, m, n, k, g, dtype, CT2B200_GEMM_AUTO, S(stream));
  });
}

CT2B200_API int ct2b200_gemm_f16(const void* a, const void* b, const void* bias, const void* residual, int act, int64_t m,
                     int64_t n, int64_t k, void* c, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    gemm_f16_tc(a, b, bias, residual, act, m, n, k, c, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_rms_norm(const void* gamma, const void* x, int64_t rows, int64_t cols, float eps, int use_residual,
                     void* y, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    launch_rms_norm(gamma, x, rows, cols, eps, use_residual != 0, y, nullptr, nullptr, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_rms_norm_quantize(const void* gamma, const void* x, int64_t rows, int64_t cols, float eps,
                              int use_residual, int8_t* q, float* scale, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(q && scale, "rms_norm_quantize: outputs are required");
    launch_rms_norm(gamma, x, rows, cols, eps, use_residual != 0, nullptr, q, scale, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_rotary(const void* x, const void* sin, const void* cos, int64_t batch, int64_t time, int64_t depth,
                   int64_t ndims, int interleave, void* y, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(ndims <= depth && ndims % 2 == 0, "rotary: ndims must be even and <= depth");
    launch_rotary(x, sin, cos, batch, time, depth, ndims, interleave != 0, y, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_softmax(const void* x, const int32_t* lengths, int64_t rows, int64_t cols, int log, void* y, int dtype,
                    void* stream) {
  return guarded([&] {
    require_device();
    launch_softmax(x, lengths, rows, cols, log != 0, y, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_topk(const void* x, int64_t rows, int64_t cols, int k, void* values, int32_t* indices, int dtype,
                 void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(k >= 1 && k <= 64 && k <= cols, "topk: k must be in [1, min(64, cols)]");
    launch_topk(x, rows, cols, k, values, indices, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_gather_rows(const void* data, const int32_t* ids, int64_t num_ids, int64_t row_bytes, void* out,
                        void* stream) {
  return guarded([&] {
    require_device();
    launch_gather_rows(data, ids, num_ids, row_bytes, out, S(stream));
  });
}

CT2B200_API int ct2b200_embedding_s8(const int8_t* w, const float* scale, const int32_t* ids, int64_t num_ids, int64_t depth,
                         void* y, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    launch_embedding_s8(w, scale, ids, num_ids, depth, y, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_mul_quantize(const void* gate, const void* up, int64_t rows, int64_t cols, int8_t* q, float* scale,
                         int dtype, void* stream) {
  return guarded([&] {
    require_device();
    launch_mul_quantize(gate, up, rows, cols, q, scale, dtype, S(stream));
  });
}

CT2B200_API size_t ct2b200_attention_decode_workspace(int64_t batch, int num_heads, int head_dim, int64_t max_len) {
  // 16 slots per (row, head) for the split-KV kernel + 64 for the persistent kernel (attention_decode.cu)
  (void)max_len;
  return attention_decode_workspace_bytes(batch, num_heads, head_dim, 80);
}

CT2B200_API int ct2b200_attention_decode(const void* qkv, void* k_cache, void* v_cache, const float* sin, const float* cos,
                             const int32_t* lens, int64_t batch, int num_heads, int num_heads_kv, int head_dim,
                             int64_t max_len, int rotary_interleave, float scale, void* out, void* workspace,
                             size_t workspace_bytes, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int splits = attention_decode_splits(batch, num_heads_kv, max_len, sms);
    launch_attention_decode(qkv, k_cache, v_cache, sin, cos, lens, batch, num_heads, num_heads_kv, head_dim, max_len,
                            rotary_interleave != 0, scale, out, workspace, workspace_bytes, splits, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_attention_prefill(const void* qkv, void* k_cache, void* v_cache, const float* sin, const float* cos,
                              const int32_t* lengths, int64_t batch, int64_t time, int64_t offset, int num_heads,
                              int num_heads_kv, int head_dim, int64_t max_len, int rotary_interleave, float scale,
                              void* out, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(offset + time <= max_len, "attention_prefill: offset + time exceeds max_len");
    launch_rope_append(const_cast<void*>(qkv), k_cache, v_cache, sin, cos, lengths, batch, time, offset, num_heads,
                       num_heads_kv, head_dim, max_len, rotary_interleave != 0, dtype, S(stream));
    launch_attention_prefill(qkv, k_cache, v_cache, lengths, batch, time, offset, num_heads, num_heads_kv, head_dim,
                             max_len, scale, out, dtype, S(stream));
  });
}

CT2B200_API int ct2b200_awq_repack(const int32_t* qweight, const void* scales, const int32_t* qzeros, int layout, int group_size,
                       int64_t n, int64_t k, int32_t* wp, void* sc, void* zr, void* sz, void* stream) {
  return guarded([&] {
    require_device();
    awq_repack(qweight, scales, qzeros, layout, group_size, n, k, wp, sc, zr, S(stream));
    if (sz) {
      AwqNative w{wp, sc, zr, n, k, group_size};
      awq_build_group_major(w, sz, S(stream));
    }
  });
}

CT2B200_API int ct2b200_dense_awq(const void* x, const int32_t* wp, const void* sc, const void* zr, const void* sz,
                      int group_size, const void* bias, const void* residual, int act, int64_t m, int64_t n, int64_t k,
                      void* y, void* scratch_nk, void* stream) {
  return guarded([&] {
    require_device();
    AwqNative w{wp, sc, zr, n, k, group_size, sz};
    dense_awq(x, w, bias, residual, act, m, y, scratch_nk, S(stream));
  });
}

CT2B200_API int ct2b200_dense_awq_glu(const void* x, const int32_t* wp_gate, const void* sc_gate, const void* zr_gate,
                          const void* sz_gate, const int32_t* wp_up, const void* sc_up, const void* zr_up,
                          const void* sz_up, int group_size, int act, int64_t m, int64_t n, int64_t k, void* h,
                          void* scratch_nk, void* scratch_mn, void* stream) {
  return guarded([&] {
    require_device();
    AwqNative g{wp_gate, sc_gate, zr_gate, n, k, group_size, sz_gate}, u{wp_up, sc_up, zr_up, n, k, group_size, sz_up};
    dense_awq_glu(x, g, u, act, m, h, scratch_nk, scratch_mn, S(stream));
  });
}

CT2B200_API int ct2b200_dequantize_awq(const int32_t* qweight, const void* scales, const int32_t* qzeros, int layout,
                           int group_size, int64_t n, int64_t k, void* w, void* stream) {
  return guarded([&] {
    require_device();
    awq_dequantize_ref_layout(qweight, scales, qzeros, layout, group_size, n, k, w, S(stream));
  });
}

// ---- engine ----
CT2B200_API ct2b200_generator* ct2b200_generator_open(const char* model_dir, const ct2b200_generator_config* config) {
  ct2b200_generator* g = nullptr;
  const int rc = guarded([&] {
    require_device();
    CT2_REQUIRE(model_dir && config, "generator_open: null argument");
    auto holder = std::make_unique<ct2b200_generator>();
    holder->impl = std::make_unique<Generator>(model_dir, *config);
    g = holder.release();
  });
  return rc == 0 ? g : nullptr;
}

CT2B200_API void ct2b200_generator_close(ct2b200_generator* g) { delete g; }

CT2B200_API int ct2b200_generator_vocab_size(const ct2b200_generator* g) {
  return g ? static_cast<int>(g->impl->decoder().config().vocab) : -1;
}

CT2B200_API int ct2b200_generator_info(const ct2b200_generator* g, int* num_layers, int* num_heads, int* num_heads_kv, int* head_dim,
                           int* d_model, int64_t* weight_bytes) {
  return guarded([&] {
    CT2_REQUIRE(g, "null generator");
    const ModelConfig& c = g->impl->decoder().config();
    if (num_layers) *num_layers = c.num_layers;
    if (num_heads) *num_heads = c.num_heads;
    if (num_heads_kv) *num_heads_kv = c.num_heads_kv;
    if (head_dim) *head_dim = c.head_dim;
    if (d_model) *d_model = static_cast<int>(c.d_model);
    if (weight_bytes) *weight_bytes = c.weight_bytes;
  });
}

CT2B200_API int ct2b200_generate_batch(ct2b200_generator* g, const int32_t* prompt_ids, const int32_t* prompt_lens, int64_t batch,
                           int64_t max_prompt_len, int64_t max_length, int64_t min_length, const int32_t* end_ids,
                           int num_end_ids, int return_end_token, int32_t* out_ids, int32_t* out_lens) {
  return guarded([&] {
    CT2_REQUIRE(g && prompt_ids && prompt_lens && out_ids && out_lens, "generate_batch: null argument");
    GenerationRequest r;
    r.prompt_ids = prompt_ids;
    r.prompt_lens = prompt_lens;
    r.batch = batch;
    r.max_prompt_len = max_prompt_len;
    r.max_length = max_length;
    r.min_length = min_length;
    r.end_ids.assign(end_ids, end_ids + (end_ids ? num_end_ids : 0));
    r.return_end_token = return_end_token != 0;
    g->impl->generate(r, out_ids, out_lens);
  });
}

CT2B200_API int ct2b200_generate_batch_scores(ct2b200_generator* g, const int32_t* prompt_ids, const int32_t* prompt_lens,
                                              int64_t batch, int64_t max_prompt_len, int64_t max_length, int64_t min_length,
                                              const int32_t* end_ids, int num_end_ids, int return_end_token,
                                              float length_penalty, int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  return guarded([&] {
    CT2_REQUIRE(g && prompt_ids && prompt_lens && out_ids && out_lens && out_scores, "generate_batch_scores: null argument");
    GenerationRequest r;
    r.prompt_ids = prompt_ids;
    r.prompt_lens = prompt_lens;
    r.batch = batch;
    r.max_prompt_len = max_prompt_len;
    r.max_length = max_length;
    r.min_length = min_length;
    r.end_ids.assign(end_ids, end_ids + (end_ids ? num_end_ids : 0));
    r.return_end_token = return_end_token != 0;
    r.return_scores = true;
    r.length_penalty = length_penalty;
    g->impl->generate(r, out_ids, out_lens, out_scores);
  });
}

CT2B200_API int ct2b200_generate_batch_beam(ct2b200_generator* g, const int32_t* prompt_ids, int64_t batch, int64_t prompt_len,
                                int64_t max_length, int64_t min_length, const int32_t* end_ids, int num_end_ids,
                                int return_end_token, int beam_size, float patience, float length_penalty, int num_hypotheses,
                                int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  return guarded([&] {
    CT2_REQUIRE(g && prompt_ids && out_ids && out_lens && out_scores, "generate_batch_beam: null argument");
    std::vector<int32_t> lens(static_cast<size_t>(std::max<int64_t>(batch, 0)), static_cast<int32_t>(prompt_len));
    GenerationRequest r;
    r.prompt_ids = prompt_ids;
    r.prompt_lens = lens.data();
    r.batch = batch;
    r.max_prompt_len = prompt_len;
    r.max_length = max_length;
    r.min_length = min_length;
    r.end_ids.assign(end_ids, end_ids + (end_ids ? num_end_ids : 0));
    r.return_end_token = return_end_token != 0;
    r.beam_size = beam_size;
    r.patience = patience;
    r.length_penalty = length_penalty;
    r.num_hypotheses = num_hypotheses;
    const std::vector<TranslationHypotheses> res = g->impl->generate_beam(r);
    for (int64_t b = 0; b < batch; ++b)
      for (int h = 0; h < num_hypotheses; ++h) {
        int32_t* dst = out_ids + (b * num_hypotheses + h) * max_length;
        const bool have = h < static_cast<int>(res[b].tokens.size());
        const int64_t len = have ? static_cast<int64_t>(res[b].tokens[h].size()) : 0;
        for (int64_t i = 0; i < max_length; ++i) dst[i] = i < len ? res[b].tokens[h][i] : -1;
        out_lens[b * num_hypotheses + h] = have ? static_cast<int32_t>(len) : -1;
        out_scores[b * num_hypotheses + h] = have ? res[b].scores[h] : 0.f;
      }
  });
}

CT2B200_API int ct2b200_forward_batch(ct2b200_generator* g, const int32_t* ids, int64_t batch, int64_t time, int return_log_probs,
                          float* logits) {
  return guarded([&] {
    CT2_REQUIRE(g && ids && logits, "forward_batch: null argument");
    g->impl->forward(ids, batch, time, return_log_probs != 0, logits);
  });
}

CT2B200_API int ct2b200_bench_decode(ct2b200_generator* g, int64_t batch, int64_t prompt_len, int64_t steps, int64_t warmup,
                         float* prefill_ms, float* decode_ms, int64_t* kernel_launches) {
  return guarded([&] {
    CT2_REQUIRE(g, "null generator");
    g->impl->bench_decode(batch, prompt_len, steps, warmup, prefill_ms, decode_ms, kernel_launches);
  });
}

CT2B200_API int ct2b200_model_summary(const char* model_dir, char* json_out, size_t capacity) {
  return guarded([&] {
    CT2_REQUIRE(model_dir && json_out && capacity > 0, "model_summary: null argument");
    ModelFile file(model_dir);
    const ModelConfig mc = parse_model_config(file);
    char buf[1024];
    const int n = std::snprintf(
        buf, sizeof(buf),
        "{\"spec\": \"%s\", \"binary_version\": %u, \"revision\": %u, \"num_layers\": %d, \"num_heads\": %d, "
        "\"num_heads_kv\": %d, \"head_dim\": %d, \"d_model\": %lld, \"ffn_dim\": %lld, \"vocab_size\": %lld, "
        "\"weights\": \"%s\", \"float_type\": \"%s\", \"rotary_interleave\": %s, \"rotary_base\": %.9g, \"rotary_scaling_type\": %d, "
        "\"layer_norm_epsilon\": %.9g, \"activation\": %d}",
        file.spec_name.c_str(), file.binary_version, file.revision, mc.num_layers, mc.num_heads, mc.num_heads_kv, mc.head_dim,
        static_cast<long long>(mc.d_model), static_cast<long long>(mc.ffn_dim), static_cast<long long>(mc.vocab),
        mc.weights.c_str(), mc.float_type.c_str(), mc.rotary_interleave ? "true" : "false", static_cast<double>(mc.rotary_base),
        mc.rotary_scaling_type, static_cast<double>(mc.eps), mc.activation);
    CT2_REQUIRE(n > 0 && static_cast<size_t>(n) < capacity, "model_summary: output buffer too small");
    std::memcpy(json_out, buf, static_cast<size_t>(n) + 1);
  });
}

CT2B200_API int ct2b200_generator_tp_handle(ct2b200_generator* g, void* handle64_h) {
  return guarded([&] {
    CT2_REQUIRE(g && handle64_h, "null argument");
    g->impl->decoder().tp_handle(handle64_h);
  });
}
CT2B200_API int ct2b200_generator_tp_connect(ct2b200_generator* g, const void* handles_h, int num_handles) {
  return guarded([&] {
    CT2_REQUIRE(g && handles_h, "null argument");
    g->impl->decoder().tp_connect(handles_h, num_handles);
  });
}

// ---- encoder-decoder path ----
CT2B200_API int ct2b200_gemm_f32(const float* a, const float* b, const float* bias, const float* residual, int act, int64_t m,
                     int64_t n, int64_t k, float* c, void* stream) {
  return guarded([&] {
    require_device();
    gemm_f32(a, b, bias, residual, act, m, n, k, c, S(stream));
  });
}

CT2B200_API int ct2b200_layer_norm(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t cols, float eps, void* y,
                       int8_t* q, float* scale, int round_before_cast, int dtype, void* stream) {
  return guarded([&] {
    require_device();
    CT2_REQUIRE(y || q, "layer_norm: no output requested");
    CT2_REQUIRE(!q || scale, "layer_norm: scale_d is required with q_d");
    launch_layer_norm(x, gamma, beta, rows, cols, eps, y, q, scale, round_before_cast != 0, dtype, S(stream));
  });
}

CT2B200_API ct2b200_translator* ct2b200_translator_open(const char* model_dir, const ct2b200_generator_config* config) {
  ct2b200_translator* t = nullptr;
  const int rc = guarded([&] {
    require_device();
    CT2_REQUIRE(model_dir && config, "translator_open: null argument");
    auto holder = std::make_unique<ct2b200_translator>();
    holder->impl = std::make_unique<Translator>(model_dir, *config);
    t = holder.release();
  });
  return rc == 0 ? t : nullptr;
}

CT2B200_API void ct2b200_translator_close(ct2b200_translator* t) { delete t; }

CT2B200_API int ct2b200_translator_info(const ct2b200_translator* t, int* encoder_layers, int* decoder_layers, int* num_heads,
                            int* d_model, int* source_vocab, int* target_vocab, int64_t* weight_bytes) {
  return guarded([&] {
    CT2_REQUIRE(t, "null translator");
    const Seq2SeqConfig& c = t->impl->config();
    if (encoder_layers) *encoder_layers = c.enc_layers;
    if (decoder_layers) *decoder_layers = c.dec_layers;
    if (num_heads) *num_heads = c.num_heads;
    if (d_model) *d_model = static_cast<int>(c.d_model);
    if (source_vocab) *source_vocab = static_cast<int>(c.src_vocab);
    if (target_vocab) *target_vocab = static_cast<int>(c.tgt_vocab);
    if (weight_bytes) *weight_bytes = c.weight_bytes;
  });
}

CT2B200_API int ct2b200_translator_summary(const char* model_dir, char* json_out, size_t capacity) {
  return guarded([&] {
    CT2_REQUIRE(model_dir && json_out && capacity > 0, "translator_summary: null argument");
    ModelFile file(model_dir);
    const Seq2SeqConfig mc = parse_seq2seq_config(file);
    char buf[1024];
    const int n = std::snprintf(
        buf, sizeof(buf),
        "{\"spec\": \"%s\", \"binary_version\": %u, \"revision\": %u, \"encoder_layers\": %d, \"decoder_layers\": %d, "
        "\"num_heads\": %d, \"head_dim\": %d, \"d_model\": %lld, \"ffn_dim\": %lld, \"source_vocab\": %lld, "
        "\"target_vocab\": %lld, \"weights\": \"%s\", \"pre_norm\": %s, \"activation\": %d, \"embeddings_scale\": %.9g, "
        "\"layer_norm_epsilon\": %.9g, \"round_before_cast\": %s}",
        file.spec_name.c_str(), file.binary_version, file.revision, mc.enc_layers, mc.dec_layers, mc.num_heads, mc.head_dim,
        static_cast<long long>(mc.d_model), static_cast<long long>(mc.ffn_dim), static_cast<long long>(mc.src_vocab),
        static_cast<long long>(mc.tgt_vocab), mc.weights.c_str(), mc.dec_pre_norm ? "true" : "false", mc.dec_activation,
        static_cast<double>(mc.dec_emb_scale), static_cast<double>(mc.eps), mc.round_before_cast ? "true" : "false");
    CT2_REQUIRE(n > 0 && static_cast<size_t>(n) < capacity, "translator_summary: output buffer too small");
    std::memcpy(json_out, buf, static_cast<size_t>(n) + 1);
  });
}

CT2B200_API int ct2b200_translate_batch(ct2b200_translator* t, const int32_t* source_ids, const int32_t* source_lens, int64_t batch,
                            int64_t max_source_len, int beam_size, float patience, float length_penalty,
                            int64_t max_decoding_length, int64_t min_decoding_length, int num_hypotheses, int32_t start_id,
                            const int32_t* end_ids, int num_end_ids, int return_end_token, int32_t* out_ids, int32_t* out_lens,
                            float* out_scores) {
  return guarded([&] {
    CT2_REQUIRE(t && source_ids && source_lens && out_ids && out_lens && out_scores, "translate_batch: null argument");
    TranslationRequest r;
    r.source_ids = source_ids;
    r.source_lens = source_lens;
    r.batch = batch;
    r.max_source_len = max_source_len;
    r.beam_size = beam_size;
    r.patience = patience;
    r.length_penalty = length_penalty;
    r.max_decoding_length = max_decoding_length;
    r.min_decoding_length = min_decoding_length;
    r.num_hypotheses = num_hypotheses;
    r.start_id = start_id;
    r.end_ids.assign(end_ids, end_ids + (end_ids ? num_end_ids : 0));
    r.return_end_token = return_end_token != 0;
    const std::vector<TranslationHypotheses> res = t->impl->translate(r);
    for (int64_t b = 0; b < batch; ++b)
      for (int h = 0; h < num_hypotheses; ++h) {
        int32_t* dst = out_ids + (b * num_hypotheses + h) * max_decoding_length;
        const bool have = h < static_cast<int>(res[b].tokens.size());
        const int64_t len = have ? static_cast<int64_t>(res[b].tokens[h].size()) : 0;
        for (int64_t i = 0; i < max_decoding_length; ++i) dst[i] = i < len ? res[b].tokens[h][i] : -1;
        out_lens[b * num_hypotheses + h] = have ? static_cast<int32_t>(len) : -1;
        out_scores[b * num_hypotheses + h] = have ? res[b].scores[h] : 0.f;
      }
  });
}

CT2B200_API int ct2b200_beam_decide_host(int beam_size, const int32_t* words, const int32_t* end_ids, int num_end_ids, int step,
                             int max_steps, int max_hyp, int max_candidates, int num_hypotheses, int early_exit, int include_eos,
                             int32_t* state_io, int32_t* active, int32_t* hyp_slot, int32_t* hyp_len) {
  return guarded([&] {
    CT2_REQUIRE(words && state_io && active && hyp_slot && hyp_len, "beam_decide_host: null argument");
    CT2_REQUIRE(beam_size >= 1 && beam_size <= kMaxBeam, "beam_size must be in [1, 32]");
    int w[2 * kMaxBeam];
    for (int i = 0; i < 2 * beam_size; ++i) w[i] = words[i];
    BeamDecision d;
    beam_decide(beam_size, w, end_ids, end_ids ? num_end_ids : 0, step, max_steps, state_io[2] != 0, state_io[1], state_io[0], max_hyp,
                max_candidates, num_hypotheses, early_exit, include_eos, d);
    if (!state_io[2]) {
      state_io[0] = d.num_hyp;
      state_io[1] = d.top_done;
      state_io[2] = d.finished;
    }
    for (int k = 0; k < beam_size; ++k) {
      active[k] = d.active[k];
      hyp_slot[k] = d.hyp_slot[k];
      hyp_len[k] = d.hyp_len[k];
    }
  });
}

CT2B200_API int ct2b200_translator_encode(ct2b200_translator* t, const int32_t* source_ids, const int32_t* source_lens, int64_t batch,
                              int64_t max_source_len, float* memory) {
  return guarded([&] {
    CT2_REQUIRE(t && source_ids && source_lens && memory, "translator_encode: null argument");
    t->impl->encode(source_ids, source_lens, batch, max_source_len, memory);
  });
}

CT2B200_API int ct2b200_bench_translate(ct2b200_translator* t, int64_t batch, int64_t source_len, int beam_size, int64_t steps,
                            int64_t warmup, float* encode_ms, float* decode_ms, int64_t* kernel_launches) {
  return guarded([&] {
    CT2_REQUIRE(t && encode_ms && decode_ms && kernel_launches, "bench_translate: null argument");
    t->impl->bench(batch, source_len, beam_size, steps, warmup, encode_ms, decode_ms, kernel_launches);
  });
}

// ---- Whisper ----
CT2B200_API int ct2b200_whisper_info(const ct2b200_translator* t, int* n_mels, int* max_frames, int* d_model, int* vocab_size) {
  return guarded([&] {
    CT2_REQUIRE(t, "null translator");
    const Seq2SeqConfig& c = t->impl->config();
    CT2_REQUIRE(c.whisper, "not a Whisper model");
    if (n_mels) *n_mels = static_cast<int>(c.n_mels);
    if (max_frames) *max_frames = static_cast<int>(c.max_frames);
    if (d_model) *d_model = static_cast<int>(c.d_model);
    if (vocab_size) *vocab_size = static_cast<int>(c.tgt_vocab);
  });
}

CT2B200_API int ct2b200_whisper_encode(ct2b200_translator* t, const float* features, int64_t batch, int64_t frames, float* memory) {
  return guarded([&] {
    CT2_REQUIRE(t && features && memory, "whisper_encode: null argument");
    t->impl->whisper_encode(features, batch, frames, memory);
  });
}

CT2B200_API int ct2b200_whisper_generate(ct2b200_translator* t, const float* features, int64_t batch, int64_t frames,
                             const int32_t* prompts, int64_t prompt_len, int beam_size, float patience, float length_penalty,
                             int64_t max_length, int num_hypotheses, const int32_t* suppress_ids, int num_suppress,
                             const int32_t* suppress_begin, int num_begin, int32_t sot_id, int32_t eot_id, int32_t no_speech_id,
                             int32_t no_timestamps_id, int max_initial_timestamp_index, int32_t* out_ids, int32_t* out_lens,
                             float* out_scores, float* no_speech) {
  return guarded([&] {
    CT2_REQUIRE(t && features && prompts && out_ids && out_lens && out_scores, "whisper_generate: null argument");
    WhisperRequest r;
    r.features = features;
    r.batch = batch;
    r.frames = frames;
    r.prompts = prompts;
    r.prompt_len = prompt_len;
    r.beam_size = beam_size;
    r.patience = patience;
    r.length_penalty = length_penalty;
    r.max_length = max_length;
    r.num_hypotheses = num_hypotheses;
    r.suppress_ids.assign(suppress_ids, suppress_ids + (suppress_ids ? num_suppress : 0));
    r.suppress_ids_begin.assign(suppress_begin, suppress_begin + (suppress_begin ? num_begin : 0));
    r.sot_id = sot_id;
    r.eot_id = eot_id;
    r.no_speech_id = no_speech_id;
    r.no_timestamps_id = no_timestamps_id;
    r.max_initial_timestamp_index = max_initial_timestamp_index;
    const std::vector<TranslationHypotheses> res = t->impl->whisper_generate(r, no_speech);
    for (int64_t b = 0; b < batch; ++b)
      for (int h = 0; h < num_hypotheses; ++h) {
        int32_t* dst = out_ids + (b * num_hypotheses + h) * max_length;
        const bool have = h < static_cast<int>(res[b].tokens.size());
        const int64_t len = have ? static_cast<int64_t>(res[b].tokens[h].size()) : 0;
        for (int64_t i = 0; i < max_length; ++i) dst[i] = i < len ? res[b].tokens[h][i] : -1;
        out_lens[b * num_hypotheses + h] = have ? static_cast<int32_t>(len) : -1;
        out_scores[b * num_hypotheses + h] = have ? res[b].scores[h] : 0.f;
      }
  });
}

}  // extern "C"

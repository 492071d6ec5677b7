/* ct2b200.h — C-ABI of the B200-native (sm_100a) quantized-transformer decode path.
 *
 * This is the drop-in boundary.  CTranslate2 has no C plugin interface: its
 * boundary is the set of C++ `<Device::CUDA>` template specialisations listed below.  Every entry
 * point here is what one of those specialisations would call; the comment above each function cites
 * the reference interface it replaces (paths relative to the reference tree).  INTEGRATION.md shows
 * the reference-side shims.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *  - `*_d` pointers are DEVICE pointers, row-major, caller-owned, 16-byte aligned; `*_h` are HOST;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); op-level functions
 *    never allocate and never synchronise (split-K scratch comes from ct2b200_workspace_*);
 *  - every function returns 0 on success, non-zero on error; ct2b200_last_error() gives the
 *    message (thread-local).  Shape/argument errors mirror the reference's std::invalid_argument,
 *    CUDA failures its std::runtime_error (src/cuda/utils.h:51-96);
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with an error.
 */
#ifndef CT2B200_H_
#define CT2B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CT2B200_API __attribute__((visibility("default")))
#else
#define CT2B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* include/ctranslate2/types.h:16-24 (DataType) — the float types activations may use. */
typedef enum { CT2B200_F32 = 0, CT2B200_F16 = 1, CT2B200_BF16 = 2 } ct2b200_dtype;

/* include/ctranslate2/ops/activation.h:9-17 (ActivationType), same order; -1 = none. */
typedef enum {
  CT2B200_ACT_NONE = -1, CT2B200_ACT_RELU = 0, CT2B200_ACT_GELU_TANH = 1, CT2B200_ACT_SWISH = 2,
  CT2B200_ACT_GELU = 3, CT2B200_ACT_GELU_SIGMOID = 4, CT2B200_ACT_TANH = 5, CT2B200_ACT_SIGMOID = 6
} ct2b200_activation;

/* INT8 GEMM implementation selector (diagnostics / tests; AUTO is what the engine uses). */
typedef enum { CT2B200_GEMM_AUTO = 0, CT2B200_GEMM_TCGEN05 = 1, CT2B200_GEMM_MMA_SYNC = 2 } ct2b200_gemm_impl;

CT2B200_API const char* ct2b200_last_error(void);
CT2B200_API const char* ct2b200_version(void);
/* Number of CUDA kernels this library has launched in the calling process (all threads). */
CT2B200_API int64_t ct2b200_kernel_launch_count(void);
/* Device properties the host side sizes grids with; fails when there is no sm_100 device. */
CT2B200_API int ct2b200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---------------------------------------------------------------------------------------------
 * Op level
 * ------------------------------------------------------------------------------------------- */

/* ops::Quantize::quantize<Device::CUDA,T,int8_t> — include/ctranslate2/ops/quantize.h:22-24,
 * src/ops/quantize_gpu.cu:57-105.  x [rows,cols] T -> q int8 [rows,cols], scale f32 [rows]. */
CT2B200_API int ct2b200_quantize_rows(const void* x_d, int dtype, int64_t rows, int64_t cols, int round_before_cast,
                          int8_t* q_d, float* scale_d, void* stream);

/* primitives<Device::CUDA>::gemm<int8_t,int32_t>(…trans_b=true, alpha=1, beta=0) —
 * include/ctranslate2/primitives.h:213-232, src/cuda/primitives.cu:571-597.
 * a [m,k] int8, b [n,k] int8 -> c [m,n] int32 (exact). k % 16 == 0. */
CT2B200_API int ct2b200_gemm_s8(const int8_t* a_d, const int8_t* b_d, int64_t m, int64_t n, int64_t k, int32_t* c_d,
                    int impl, void* stream);

/* ops::Dequantize::dequantize_gemm_output<Device::CUDA,T> — include/ctranslate2/ops/dequantize.h:19-25,
 * src/ops/dequantize_gpu.cu:30-144.  y = act(c / (a_scale[i]*b_scale[j]) + bias[j]).  bias may be NULL. */
CT2B200_API int ct2b200_dequantize_gemm_output(const int32_t* c_d, const float* a_scale_d, const float* b_scale_d,
                                   const void* bias_d, int act, int64_t m, int64_t n, void* y_d, int dtype,
                                   void* stream);

/* ops::Dequantize::dequantize<Device::CUDA,int8_t,T> (embedding rows) — src/ops/dequantize_gpu.cu:16-27:
 * y[i,:] = x[i,:] / scale[i]. */
CT2B200_API int ct2b200_dequantize_rows(const int8_t* x_d, const float* scale_d, int64_t rows, int64_t cols, void* y_d,
                            int dtype, void* stream);

/* layers::Dense::operator(), quantized arm, as ONE fused launch — src/layers/common.cc:353-401:
 *   y = act(gemm_s8(xq, w) / (x_scale[i]*w_scale[j]) + bias[j]) + residual[i,j]
 * xq [m,k] int8 with x_scale [m] (from ct2b200_quantize_rows / ct2b200_rms_norm_quantize),
 * w [n,k] int8 with w_scale [n]; bias [n] T or NULL; residual [m,n] T or NULL; y [m,n] T. */
CT2B200_API int ct2b200_dense_s8(const int8_t* xq_d, const float* x_scale_d, const int8_t* w_d, const float* w_scale_d,
                     const void* bias_d, const void* residual_d, int act, int64_t m, int64_t n, int64_t k,
                     void* y_d, int dtype, int impl, void* stream);

/* FeedForwardNetwork gate/up pair (src/layers/transformer.cc:21-51 with ffn_glu): fused
 *   h = act(dense(xq, w_gate)) * dense(xq, w_up)       h [m,n] T
 * replacing Dense(linear_0)+Dense(linear_0_noact)+ops::Mul. */
CT2B200_API int ct2b200_dense_s8_glu(const int8_t* xq_d, const float* x_scale_d, const int8_t* w_gate_d,
                         const float* w_gate_scale_d, const int8_t* w_up_d, const float* w_up_scale_d, int act,
                         int64_t m, int64_t n, int64_t k, void* h_d, int dtype, int impl, void* stream);

/* The whole quantized arm of layers::Dense::operator() INCLUDING its input side —
 * src/layers/common.cc:353-401 preceded by ops::Quantize (quantize.cc:21-50) or, with gamma_d, by the layer's pre-norm
 * ops::RMSNorm (rms_norm_gpu.cu:19-63):  xq, x_scale = Quantize([RMSNorm(x, gamma, eps)]);  y = dense_s8(xq, x_scale, ...).
 * x [m,k] T; xq_d [m,k] int8 and x_scale_d [m] are OUTPUTS (the same bits ct2b200_quantize_rows / ct2b200_rms_norm_quantize
 * produce).  Runs as the register-resident row kernel + the fused Dense under programmatic dependent launch; barrier_d is
 * unused (a variant that ran the row op inside the GEMM behind a grid barrier measured slower on the B200 and was removed). */
CT2B200_API int ct2b200_dense_s8_rows(const void* x_d, const void* gamma_d, float eps, const int8_t* w_d, const float* w_scale_d,
                          const void* bias_d, const void* residual_d, int act, int64_t m, int64_t n, int64_t k,
                          void* y_d, int dtype, int8_t* xq_d, float* x_scale_d, unsigned* barrier_d, void* stream);
/* same for the gate/up pair: h = act(dense(xq, w_gate)) * dense(xq, w_up) */
CT2B200_API int ct2b200_dense_s8_glu_rows(const void* x_d, const void* gamma_d, float eps, const int8_t* w_gate_d,
                              const float* w_gate_scale_d, const int8_t* w_up_d, const float* w_up_scale_d, int act,
                              int64_t m, int64_t n, int64_t k, void* h_d, int dtype, int8_t* xq_d, float* x_scale_d,
                              unsigned* barrier_d, void* stream);

/* primitives<Device::CUDA>::gemm<float16_t|bfloat16_t> (trans_b, alpha 1, beta 0) + ops::Gemm's
 * apply_bias_and_activation — src/cuda/primitives.cu:485-569, src/ops/gemm.cc:10-25.
 * a [m,k] T, b [n,k] T -> c [m,n] T, fp32 accumulation; dtype F16 or BF16. */
CT2B200_API int ct2b200_gemm_f16(const void* a_d, const void* b_d, const void* bias_d, const void* residual_d, int act,
                     int64_t m, int64_t n, int64_t k, void* c_d, int dtype, void* stream);

/* primitives<Device::CUDA>::gemm<float,float> (trans_b, alpha 1, beta 0) + apply_bias_and_activation —
 * src/cuda/primitives.cu:485-505 (cublasSgemm), src/ops/gemm.cc:10-25.  True fp32 FMAs (no TF32). */
CT2B200_API int ct2b200_gemm_f32(const float* a_d, const float* b_d, const float* bias_d, const float* residual_d, int act,
                     int64_t m, int64_t n, int64_t k, float* c_d, void* stream);

/* ops::LayerNorm::compute<Device::CUDA,T> (last axis) — include/ctranslate2/ops/layer_norm.h, src/ops/layer_norm_gpu.cu:33-66,
 * 169-206: y = (x - mean) * rsqrt(var + eps) * gamma + beta.  y_d may be NULL when only the quantized row is wanted;
 * q_d / scale_d non-NULL adds ops::Quantize of T(y) in the same launch (layers::LayerNorm + Dense's Quantize). */
CT2B200_API int ct2b200_layer_norm(const void* x_d, const void* gamma_d, const void* beta_d, int64_t rows, int64_t cols, float eps,
                       void* y_d, int8_t* q_d, float* scale_d, int round_before_cast, int dtype, void* stream);

/* ops::RMSNorm::compute<Device::CUDA,T> — include/ctranslate2/ops/rms_norm.h, src/ops/rms_norm_gpu.cu:19-63. */
CT2B200_API int ct2b200_rms_norm(const void* gamma_d, const void* x_d, int64_t rows, int64_t cols, float eps,
                     int use_residual, void* y_d, int dtype, void* stream);

/* RMSNorm followed by Quantize in one launch (layers::LayerNorm + Dense's Quantize,
 * src/layers/common.cc:464-472 + :392): q = quantize(T(rms_norm(x))). */
CT2B200_API int ct2b200_rms_norm_quantize(const void* gamma_d, const void* x_d, int64_t rows, int64_t cols, float eps,
                              int use_residual, int8_t* q_d, float* scale_d, int dtype, void* stream);

/* ops::Rotary::compute<Device::CUDA,T> — include/ctranslate2/ops/rotary.h, src/ops/rotary_gpu.cu:27-85.
 * x [batch, time, depth] rows (batch = b*h when transposed), sin/cos [time, ndims] T. */
CT2B200_API int ct2b200_rotary(const void* x_d, const void* sin_d, const void* cos_d, int64_t batch, int64_t time,
                   int64_t depth, int64_t ndims, int interleave, void* y_d, int dtype, void* stream);

/* ops::SoftMax::compute<Device::CUDA,T> (log=0) / LogSoftMax (log=1) — src/ops/softmax_gpu.cu:190-256.
 * lengths_d int32 [rows] or NULL. */
CT2B200_API int ct2b200_softmax(const void* x_d, const int32_t* lengths_d, int64_t rows, int64_t cols, int log, void* y_d,
                    int dtype, void* stream);

/* ops::TopK::compute<Device::CUDA,T,int32_t> — include/ctranslate2/ops/topk.h, src/ops/topk_gpu.cu:181-335.
 * Descending values; exact ties resolve lowest index first.  k <= 64. */
CT2B200_API int ct2b200_topk(const void* x_d, int64_t rows, int64_t cols, int k, void* values_d, int32_t* indices_d,
                 int dtype, void* stream);

/* ops::Gather::compute<Device::CUDA,T>(axis 0) — src/ops/gather_gpu.cu:52-91.  Row copy of `row_bytes`. */
CT2B200_API int ct2b200_gather_rows(const void* data_d, const int32_t* ids_d, int64_t num_ids, int64_t row_bytes,
                        void* out_d, void* stream);

/* layers::Embeddings::operator() with INT8 weights — src/layers/common.cc:64-81
 * (Gather rows + Gather scales + Dequantize) in one launch: y[i,:] = w[ids[i],:] / scale[ids[i]]. */
CT2B200_API int ct2b200_embedding_s8(const int8_t* w_d, const float* scale_d, const int32_t* ids_d, int64_t num_ids,
                         int64_t depth, void* y_d, int dtype, void* stream);

/* ops::Mul + ops::Quantize of the SwiGLU product — q = quantize(T(gate*up)) (transformer.cc:31-37). */
CT2B200_API int ct2b200_mul_quantize(const void* gate_d, const void* up_d, int64_t rows, int64_t cols, int8_t* q_d,
                         float* scale_d, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention: layers::MultiHeadAttention::operator() between the QKV Dense and
 * the output Dense — src/layers/attention.cc:485-602; FlashMultiHeadAttention src/layers/flash_attention.cc:18-137.
 *
 * KV cache layout (un-replicated GQA): k_cache / v_cache [batch_slots, num_heads_kv, max_len, head_dim] T.
 * qkv [rows, (H + 2*Hkv) * head_dim] T is the fused linear_0 output ([q | k | v]).
 * sin/cos tables [max_positions, head_dim] f32 (built like RotaryEmbeddings::initialize).
 * ------------------------------------------------------------------------------------------- */

/* Decode step (one new token per sequence): rotary(q,k) at position lens[b], append k/v at lens[b],
 * softmax(q k^T / sqrt(d)) v over positions 0..lens[b].  out [batch, H*head_dim] T.
 * lens_d int32 [batch] = tokens already cached per row (not modified).  head_dim must be 128 or 64 or 32.
 * The caches must hold finite values everywhere (zero-initialise them once): whole 64-key boxes are staged and the
 * keys past lens_d[b] are masked, not skipped.  workspace_d: ct2b200_attention_decode_workspace bytes, zeroed once. */
CT2B200_API int ct2b200_attention_decode(const void* qkv_d, void* k_cache_d, void* v_cache_d, const float* sin_d,
                             const float* cos_d, const int32_t* lens_d, int64_t batch, int num_heads,
                             int num_heads_kv, int head_dim, int64_t max_len, int rotary_interleave,
                             float scale, void* out_d, void* workspace_d, size_t workspace_bytes, int dtype,
                             void* stream);
CT2B200_API size_t ct2b200_attention_decode_workspace(int64_t batch, int num_heads, int head_dim, int64_t max_len);

/* Prefill (T new tokens per sequence starting at position `offset`, causal): qkv [batch*time, ...],
 * lengths_d int32 [batch] = valid new tokens per row (NULL = time).  out [batch*time, H*head_dim] T. */
CT2B200_API int ct2b200_attention_prefill(const void* qkv_d, void* k_cache_d, void* v_cache_d, const float* sin_d,
                              const float* cos_d, const int32_t* lengths_d, int64_t batch, int64_t time,
                              int64_t offset, int num_heads, int num_heads_kv, int head_dim, int64_t max_len,
                              int rotary_interleave, float scale, void* out_d, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AWQ-INT4: ops::GemmAwq / GemvAwq / DequantizeAwq — include/ctranslate2/ops/awq/{gemm,gemv,dequantize}.h,
 * src/ops/awq/{gemm,gemv,dequantize}_gpu.cu.  x [m,k] f16 -> y [m,n] f16.
 * layout 1 = AWQ_GEMM: qweight int32 [k, n/8], scales f16 [k/g, n], qzeros int32 [k/g, n/8]
 * layout 2 = AWQ_GEMV: qweight int32 [n, k/8], scales f16 [n, sf_w], qzeros int32 [n, zeros_w]
 * ------------------------------------------------------------------------------------------- */
/* One-time repack (Dense ctor / model load) of either reference layout into the native K-major layout:
 * wp int32 [n, k/8] (channel 8w+i of row n in nibble {0,4,1,5,2,6,3,7}[i] of word w), sc f16 [n, k/g], zr f16 [n, k/g],
 * and (sz_d non-NULL) the same {scale, zero} values as f16 pairs in group-major order sz [k/g, n][2]: the decode kernel
 * (m <= 64) stages the 128 pairs of a tile and group with one bulk copy and needs it; NULL = not produced. */
CT2B200_API int ct2b200_awq_repack(const int32_t* qweight_d, const void* scales_d, const int32_t* qzeros_d, int layout,
                       int group_size, int64_t n, int64_t k, int32_t* wp_d, void* sc_d, void* zr_d, void* sz_d,
                       void* stream);

/* ops::GemmAwq / GemvAwq + apply_bias_and_activation — src/ops/awq/gemm.cc:8-33, gemv.cc:9-37, on the native layout:
 * y = act(x . deq(W)^T + bias) + residual.  m <= 64: fused dequantize + tcgen05 GEMM (weight-streaming kernel with the
 * operand in tensor memory when sz_d is given, the general kernel otherwise).  m > 64 (the reference's
 * DequantizeAwq + cuBLAS arm, src/layers/common.cc:409-420): needs scratch_nk_d, an fp16 [n,k] buffer. */
CT2B200_API int ct2b200_dense_awq(const void* x_d, const int32_t* wp_d, const void* sc_d, const void* zr_d, const void* sz_d,
                      int group_size, const void* bias_d, const void* residual_d, int act, int64_t m, int64_t n, int64_t k,
                      void* y_d, void* scratch_nk_d, void* stream);
/* gate/up pair of the gated FFN in one pass: h = act(x . deq(Wg)^T) * (x . deq(Wu)^T).  m > 64 also needs scratch_mn_d. */
CT2B200_API int ct2b200_dense_awq_glu(const void* x_d, const int32_t* wp_gate_d, const void* sc_gate_d, const void* zr_gate_d,
                          const void* sz_gate_d, const int32_t* wp_up_d, const void* sc_up_d, const void* zr_up_d,
                          const void* sz_up_d, int group_size, int act, int64_t m, int64_t n, int64_t k, void* h_d,
                          void* scratch_nk_d, void* scratch_mn_d, void* stream);
/* ops::DequantizeAwq — src/ops/awq/dequantize_gpu.cu:8-62: reference layout (1 or 2) -> W f16 [k, n]. */
CT2B200_API int ct2b200_dequantize_awq(const int32_t* qweight_d, const void* scales_d, const int32_t* qzeros_d, int layout,
                           int group_size, int64_t n, int64_t k, void* w_d /* f16 [k,n] */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Engine level: models::Model::load + Generator (§3.1).
 * ctranslate2::Generator — include/ctranslate2/generator.h:11-39; GenerationOptions generation.h:14-78.
 * Token strings <-> ids (Vocabulary) stay on the caller's side of the boundary; ids cross it.
 * ------------------------------------------------------------------------------------------- */
typedef struct ct2b200_generator ct2b200_generator;

/* Weight type requested by the compute type (models::Model::set_compute_type / ensure_dtype, src/models/model.cc:178-234,
 * 304-369): STORED = "default" (keep what model.bin holds), INT8 = the int8* compute types (float weights are quantized at
 * load: scale = 127 / amax per row, q = rint(w * scale)), FLOAT = float16 / bfloat16 (int8 weights are dequantized at load).
 * The conversion runs on the GPU.  AWQ-INT4 models ignore it (the reference pins ComputeType::FLOAT16, model.cc:750-757). */
typedef enum { CT2B200_WEIGHTS_STORED = 0, CT2B200_WEIGHTS_INT8 = 1, CT2B200_WEIGHTS_FLOAT = 2 } ct2b200_weight_type;

typedef struct {
  int device;               /* CUDA device ordinal */
  int compute_type;         /* ct2b200_dtype of activations / KV cache */
  int64_t max_batch;        /* batch slots to reserve */
  int64_t max_length;       /* max total positions (prompt + generated) per sequence */
  int tp_rank, tp_size;     /* tensor-parallel rank/size (1 = off); see ct2b200_generator_tp_connect */
  int use_cuda_graph;       /* capture the decode step in a CUDA graph */
  int gemm_impl;            /* ct2b200_gemm_impl */
  int weight_type;          /* ct2b200_weight_type: what Model::set_compute_type asks of the Dense / embedding weights */
} ct2b200_generator_config;

/* models::Model::load(model_dir, Device::CUDA, device, compute_type) + Generator ctor.
 * Reads model.bin (binary versions 2..6), config.json. */
CT2B200_API ct2b200_generator* ct2b200_generator_open(const char* model_dir, const ct2b200_generator_config* config);
CT2B200_API void ct2b200_generator_close(ct2b200_generator* g);
CT2B200_API int ct2b200_generator_vocab_size(const ct2b200_generator* g);
CT2B200_API int ct2b200_generator_info(const ct2b200_generator* g, int* num_layers, int* num_heads, int* num_heads_kv,
                           int* head_dim, int* d_model, int64_t* weight_bytes);

/* Host only (no device needed): what models::Model::load would find in `model_dir` — spec, binary version, decoder geometry
 * (layers, heads, kv heads, head_dim, d_model, ffn_dim, vocabulary) and the storage type of the linear layers, as a JSON
 * object written to json_out.  The same parser configures ct2b200_generator_open. */
CT2B200_API int ct2b200_model_summary(const char* model_dir, char* json_out, size_t capacity);

/* Generator::generate_batch_async(...).get(), greedy (beam_size 1, sampling_topk 1),
 * include_prompt_in_result=false.  HOST buffers:
 *   prompt_ids_h [batch, max_prompt_len] int32 (right-padded), prompt_lens_h [batch];
 *   end_ids_h [num_end_ids]; out_ids_h [batch, max_length] int32 (filled with -1 past the end),
 *   out_lens_h [batch].  max_length / min_length count generated tokens (generation.h:41-43). */
CT2B200_API int ct2b200_generate_batch(ct2b200_generator* g, const int32_t* prompt_ids_h, const int32_t* prompt_lens_h,
                           int64_t batch, int64_t max_prompt_len, int64_t max_length, int64_t min_length,
                           const int32_t* end_ids_h, int num_end_ids, int return_end_token,
                           int32_t* out_ids_h, int32_t* out_lens_h);

/* The same with GenerationOptions::return_scores = true: out_scores_h [batch] = sum of the log-probabilities of the
 * generated tokens (LogSoftMax of the processed logits, the end token's included) / length^length_penalty
 * (src/decoding.cc:875-923, 189-203; include/ctranslate2/generation.h:22-23, 55). */
CT2B200_API int ct2b200_generate_batch_scores(ct2b200_generator* g, const int32_t* prompt_ids_h, const int32_t* prompt_lens_h,
                           int64_t batch, int64_t max_prompt_len, int64_t max_length, int64_t min_length,
                           const int32_t* end_ids_h, int num_end_ids, int return_end_token, float length_penalty,
                           int32_t* out_ids_h, int32_t* out_lens_h, float* out_scores_h);

/* Generator::generate_batch_async with beam_size > 1 (BeamSearch::search, src/decoding.cc:425-720; GenerationOptions beam_size,
 * patience, length_penalty, num_hypotheses): prompts of equal length; out_ids_h [batch, num_hypotheses, max_length] (-1 padded),
 * out_lens_h / out_scores_h [batch, num_hypotheses] (length -1 = fewer hypotheses than asked); batch * beam_size <= max_batch.
 * Scores: cumulative log-probability / length^length_penalty, the end token counted (decoding.h:154). */
CT2B200_API int ct2b200_generate_batch_beam(ct2b200_generator* g, const int32_t* prompt_ids_h, int64_t batch, int64_t prompt_len,
                                int64_t max_length, int64_t min_length, const int32_t* end_ids_h, int num_end_ids,
                                int return_end_token, int beam_size, float patience, float length_penalty, int num_hypotheses,
                                int32_t* out_ids_h, int32_t* out_lens_h, float* out_scores_h);

/* Generator::forward_batch_async(ids, return_log_probs) — full-sequence forward from position 0.
 * ids_h [batch, time] int32 host; logits_h [batch, time, vocab] f32 host. */
CT2B200_API int ct2b200_forward_batch(ct2b200_generator* g, const int32_t* ids_h, int64_t batch, int64_t time,
                          int return_log_probs, float* logits_h);

/* Split phases, device-timed, for bench.py: prefill `prompt_len-1` tokens then run `steps` decode
 * steps with inputs already resident in HBM.  Returns device milliseconds of each phase. */
CT2B200_API int ct2b200_bench_decode(ct2b200_generator* g, int64_t batch, int64_t prompt_len, int64_t steps, int64_t warmup,
                         float* prefill_ms, float* decode_ms, int64_t* kernel_launches);

/* Tensor parallel (ct2b200_generator_config.tp_size > 1; one process per GPU; replaces ScopedMPISetter + the NCCL
 * communicator of src/devices.cc:141-217 and ops::ReduceAll / GatherAll, src/ops/nccl_ops_gpu.cu:52-85).
 * The collectives of the decoder are fused into its kernels over NVLink peer memory, so the bootstrap only has to
 * exchange one CUDA IPC handle per rank:
 *   1. every rank opens the generator with its tp_rank / tp_size (weights are sharded on load as in
 *      src/models/model.cc:662-743: QKV and gate/up by output channel, out-proj and down-proj by input channel);
 *   2. ct2b200_generator_tp_handle writes this rank's 64-byte cudaIpcMemHandle_t to handle64_h;
 *   3. the caller all-gathers the handles (torch.distributed, MPI, files ...) and passes the tp_size handles in rank
 *      order to ct2b200_generator_tp_connect.
 * Every rank must then issue the same generate_batch / forward_batch calls with the same inputs. */
CT2B200_API int ct2b200_generator_tp_handle(ct2b200_generator* g, void* handle64_h);
CT2B200_API int ct2b200_generator_tp_connect(ct2b200_generator* g, const void* handles_h, int num_handles);

/* ---------------------------------------------------------------------------------------------
 * Encoder-decoder path: ctranslate2::Translator — include/ctranslate2/translator.h:20-60,
 * TranslationOptions include/ctranslate2/translation.h:14-98; models::TransformerModel (src/models/transformer.cc),
 * EncoderDecoderReplica::run_translation (src/models/sequence_to_sequence.cc:305-420), TransformerEncoder /
 * TransformerDecoder with cross-attention (src/layers/transformer.cc), BeamSearch::search (src/decoding.cc:425-720).
 * Serves pre- and post-norm LayerNorm Transformers with absolute (sinusoidal or stored) positions; model.bin versions 2..6.
 * ct2b200_generator_config: compute_type / weight_type / device / use_cuda_graph are honoured; max_length bounds the
 * positions reserved for sinusoidal encodings (>= 500); max_batch is a hint (arenas grow on demand).
 * ------------------------------------------------------------------------------------------- */
typedef struct ct2b200_translator ct2b200_translator;
CT2B200_API ct2b200_translator* ct2b200_translator_open(const char* model_dir, const ct2b200_generator_config* config);
CT2B200_API void ct2b200_translator_close(ct2b200_translator* t);
CT2B200_API int ct2b200_translator_info(const ct2b200_translator* t, int* encoder_layers, int* decoder_layers, int* num_heads,
                            int* d_model, int* source_vocab, int* target_vocab, int64_t* weight_bytes);
/* Host only: the geometry parse_seq2seq_config reads from `model_dir`, as JSON. */
CT2B200_API int ct2b200_translator_summary(const char* model_dir, char* json_out, size_t capacity);

/* Translator::translate_batch on ids (Vocabulary lookups stay on the caller's side).  HOST buffers:
 *   source_ids_h [batch, max_source_len] int32 right-padded (with_source_bos / eos already applied), source_lens_h [batch];
 *   out_ids_h [batch, num_hypotheses, max_decoding_length] (-1 padded), out_lens_h / out_scores_h [batch, num_hypotheses]
 *   (length -1 = fewer hypotheses than asked).  beam_size 1 = GreedySearch (identical results to the beam-of-one search).
 *   Scores are cumulative log-probabilities / length^length_penalty (finalize_result, decoding.cc:189-254). */
CT2B200_API int ct2b200_translate_batch(ct2b200_translator* t, const int32_t* source_ids_h, const int32_t* source_lens_h,
                            int64_t batch, int64_t max_source_len, int beam_size, float patience, float length_penalty,
                            int64_t max_decoding_length, int64_t min_decoding_length, int num_hypotheses, int32_t start_id,
                            const int32_t* end_ids_h, int num_end_ids, int return_end_token, int32_t* out_ids_h,
                            int32_t* out_lens_h, float* out_scores_h);
/* Host only (no device): the per-entry bookkeeping of one BeamSearch::search step (decoding.cc:595-663) — the SAME function the
 * device kernel runs (csrc/kernels/beam_decide.h).  words_h [2 * beam_size] = the candidates' tokens in TopK order.  Outputs:
 * active_h [beam] (candidate each next beam continues), hyp_slot_h / hyp_len_h [beam] (hypothesis registered for candidate k, or
 * -1), state_io_h = {num_hyp, top_done, finished} updated in place. */
CT2B200_API int ct2b200_beam_decide_host(int beam_size, const int32_t* words_h, const int32_t* end_ids_h, int num_end_ids, int step,
                             int max_steps, int max_hyp, int max_candidates, int num_hypotheses, int early_exit,
                             int include_eos, int32_t* state_io_h, int32_t* active_h, int32_t* hyp_slot_h, int32_t* hyp_len_h);

/* TransformerEncoder::operator() — memory_h [batch, max_source_len, d_model] f32 host (padded positions are unspecified). */
CT2B200_API int ct2b200_translator_encode(ct2b200_translator* t, const int32_t* source_ids_h, const int32_t* source_lens_h,
                              int64_t batch, int64_t max_source_len, float* memory_h);
/* Device-timed phases for bench.py: one encoder pass of [batch, source_len], then `steps` beam-search steps of
 * batch * beam_size rows with inputs resident in HBM. */
CT2B200_API int ct2b200_bench_translate(ct2b200_translator* t, int64_t batch, int64_t source_len, int beam_size, int64_t steps,
                            int64_t warmup, float* encode_ms, float* decode_ms, int64_t* kernel_launches);

/* ---------------------------------------------------------------------------------------------
 * Whisper: ctranslate2::models::Whisper — include/ctranslate2/models/whisper.h:86-190, src/models/whisper.cc,
 * WhisperEncoder / WhisperDecoder src/layers/whisper.cc, ops::Conv1D src/ops/conv1d_gpu.cu.  A WhisperSpec directory opens
 * with ct2b200_translator_open (same handle type; the encoder is the Conv1D front-end instead of source embeddings).
 * Served: encode, and generate for prompts made of previous-text tokens, <|startoftranscript|> and the task tokens (no text
 * after them), with the timestamp rules (whisper.cc:742-860) unless the last task token is <|notimestamps|>;
 * detect_language / align are not provided.
 * ------------------------------------------------------------------------------------------- */
CT2B200_API int ct2b200_whisper_info(const ct2b200_translator* t, int* n_mels, int* max_frames, int* d_model, int* vocab_size);
/* Whisper::encode: features_h [batch, n_mels, frames] f32 host -> memory_h [batch, (frames + 1) / 2, d_model] f32 host. */
CT2B200_API int ct2b200_whisper_encode(ct2b200_translator* t, const float* features_h, int64_t batch, int64_t frames,
                           float* memory_h);
/* Whisper::generate(features, prompts, WhisperOptions): prompts_h [batch, prompt_len] ids; suppress_ids_h = the resolved
 * WhisperOptions::suppress_tokens (config.json "suppress_ids" for -1), suppress_begin_h = "suppress_ids_begin" when
 * suppress_blank; out_ids_h [batch, num_hypotheses, max_length] (-1 padded; the decoder runs min(max_length / 2,
 * max_length - prompt_len + 1) steps, whisper.cc:299), out_lens_h / out_scores_h [batch, num_hypotheses];
 * no_speech_h [batch] or NULL (return_no_speech_prob; needs no_speech_id). */
CT2B200_API int ct2b200_whisper_generate(ct2b200_translator* t, const float* features_h, int64_t batch, int64_t frames,
                             const int32_t* prompts_h, int64_t prompt_len, int beam_size, float patience, float length_penalty,
                             int64_t max_length, int num_hypotheses, const int32_t* suppress_ids_h, int num_suppress,
                             const int32_t* suppress_begin_h, int num_begin, int32_t sot_id, int32_t eot_id,
                             int32_t no_speech_id, int32_t no_timestamps_id, int max_initial_timestamp_index,
                             int32_t* out_ids_h, int32_t* out_lens_h, float* out_scores_h, float* no_speech_h);

#ifdef __cplusplus
}
#endif
#endif /* CT2B200_H_ */

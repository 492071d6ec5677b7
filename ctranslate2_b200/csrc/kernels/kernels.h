// kernels.h — host-callable launchers of every CUDA kernel in csrc/kernels (one stream, no allocation
// except the lazily created split-K scratch, no synchronisation).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../common.cuh"
#include "gemm_common.cuh"

namespace ct2b200 {

// rowwise.cu
void launch_quantize_rows(const void* x, int dtype, int64_t rows, int64_t cols, bool round, int8_t* q,
                          float* scale, cudaStream_t st);
void launch_rms_norm(const void* gamma, const void* x, int64_t rows, int64_t cols, float eps, bool use_residual,
                     void* y, int8_t* q, float* scale, int dtype, cudaStream_t st);
void launch_mul_quantize(const void* a, const void* b, int64_t rows, int64_t cols, int8_t* q, float* scale,
                         int dtype, cudaStream_t st);
void launch_dequantize_rows(const int8_t* x, const float* scale, int64_t rows, int64_t cols, void* y, int dtype,
                            cudaStream_t st, bool reciprocal = false);
void launch_dequantize_gemm_output(const int32_t* c, const DenseEpilogue& e, int64_t m, int64_t n, int dtype,
                                   cudaStream_t st);
void launch_embedding_s8(const int8_t* w, const float* scale, const int32_t* ids, int64_t num_ids, int64_t depth,
                         void* y, int dtype, cudaStream_t st);
void launch_gather_rows(const void* data, const int32_t* ids, int64_t num_ids, int64_t row_bytes, void* out,
                        cudaStream_t st);
void launch_rotary(const void* x, const void* sin, const void* cos, int64_t batch, int64_t time, int64_t depth,
                   int64_t ndims, bool interleave, void* y, int dtype, cudaStream_t st);
void launch_softmax(const void* x, const int32_t* lengths, int64_t rows, int64_t cols, bool log, void* y,
                    int dtype, cudaStream_t st);
void launch_topk(const void* x, int64_t rows, int64_t cols, int k, void* values, int32_t* indices, int dtype,
                 cudaStream_t st);

// tp_rows.cu — tensor-parallel row kernels (collectives fused into their consumers over NVLink peer memory)
struct TpLink {
  int rank = 0, world = 1;
  const uint32_t* tick = nullptr;             // forward-pass counter of this rank (device)
  uint32_t* flags_local = nullptr;            // [2][8] epoch flags in this rank's exchange buffer (peers write them)
  uint32_t* flags_peer[8] = {};               // the same array inside every peer's buffer
  const void* parts[2][8] = {};               // partial buffer b ([rows, d_model] T) of rank r
  unsigned long long* amax_local = nullptr;   // [2][8][amax_rows] {epoch, amax} words in this rank's buffer
  unsigned long long* amax_peer[8] = {};
  int64_t amax_rows = 0;
};
void launch_tp_tick(uint32_t* tick, cudaStream_t st);
void launch_tp_reduce_norm_quantize(const TpLink& tp, int buf, int sync_idx, void* x, const void* gamma, int64_t rows,
                                    int64_t cols, float eps, int8_t* q, float* scale, int dtype, cudaStream_t st);
void launch_tp_reduce_norm(const TpLink& tp, int buf, int sync_idx, void* x, const void* gamma, int64_t rows, int64_t cols,
                           float eps, void* y, int dtype, cudaStream_t st);
void launch_tp_reduce(const TpLink& tp, int buf, int sync_idx, void* x, int64_t rows, int64_t cols, int dtype,
                      cudaStream_t st);
void launch_tp_quantize_rows(const TpLink& tp, int slot, int sync_idx, const void* x, int64_t rows, int64_t cols, int8_t* q,
                             float* scale, int dtype, cudaStream_t st);

// gemm_tc.cu (tcgen05) — gemm_s8_mma.cu declarations live in gemm_common.cuh
void gemm_s8_tc(const int8_t* A, const int8_t* B, int64_t M, int64_t N, int64_t K, const DenseEpilogue& epi,
                int dtype, cudaStream_t st);
void gemm_s8_glu_tc(const int8_t* A, const int8_t* Bgate, const int8_t* Bup, int64_t M, int64_t N, int64_t K,
                    const GluEpilogue& glu, int dtype, cudaStream_t st);
void gemm_f16_tc(const void* A, const void* B, const void* bias, const void* residual, int act, int64_t M,
                 int64_t N, int64_t K, void* C, int dtype, cudaStream_t st);

// Row pre-phase of the decode GEMM: the kernel quantizes its own activations (ops::Quantize or ops::RMSNorm + ops::Quantize
// of x [M, K] in the GEMM's dtype, bit-identical to launch_quantize_rows / launch_rms_norm) into q / s before it stages them,
// behind a grid barrier.  q must be the `A` pointer of the call and s the epilogue's a_scale; bar = 2 zero-initialised words.
struct RowPre {
  int mode = 0;                 // 1 Quantize, 2 RMSNorm + Quantize
  const void* x = nullptr;      // [M, K] T
  const void* gamma = nullptr;  // [K] T (mode 2)
  float eps = 0.f;
  unsigned* bar = nullptr;
};
// The Dense that follows in the decode step (same m): its int8 weights are prefetched into L2 while this one streams
// (successor prefetch, gemm_decode_common.cuh).  w2 = the "up" matrix when the successor is a gate/up pair.
struct NextWeights {
  const void* w = nullptr;
  const void* w2 = nullptr;
  int64_t n = 0, k = 0;
};
// gemm_decode.cu (tcgen05, m <= 64): false = shape (or pre-phase) not covered, use the general kernel (+ separate row kernel)
bool gemm_s8_decode(const int8_t* A, const int8_t* B, int64_t M, int64_t N, int64_t K, const DenseEpilogue& epi,
                    int dtype, cudaStream_t st, const RowPre* pre = nullptr, const NextWeights* next = nullptr);
bool gemm_s8_glu_decode(const int8_t* A, const int8_t* Bgate, const int8_t* Bup, int64_t M, int64_t N, int64_t K,
                        const GluEpilogue& glu, int dtype, cudaStream_t st, const RowPre* pre = nullptr,
                        const NextWeights* next = nullptr);
bool gemm_f16_decode(const void* A, const void* B, const void* bias, const void* residual, int act, int64_t M,
                     int64_t N, int64_t K, void* C, int dtype, cudaStream_t st);

// gemm_prefill.cu (tcgen05, m > 64, compute bound): false = shape not covered
bool gemm_s8_prefill(const int8_t* A, const int8_t* B, int64_t M, int64_t N, int64_t K, const DenseEpilogue& epi,
                     int dtype, cudaStream_t st);
bool gemm_s8_glu_prefill(const int8_t* A, const int8_t* Bgate, const int8_t* Bup, int64_t M, int64_t N, int64_t K,
                         const GluEpilogue& glu, int dtype, cudaStream_t st);
bool gemm_f16_prefill(const void* A, const void* B, const void* bias, const void* residual, int act, int64_t M,
                      int64_t N, int64_t K, void* C, int dtype, cudaStream_t st);

// dispatch by ct2b200_gemm_impl
void gemm_s8(const int8_t* A, const int8_t* B, int64_t M, int64_t N, int64_t K, const DenseEpilogue& epi,
             int dtype, int impl, cudaStream_t st);
void gemm_s8_glu(const int8_t* A, const int8_t* Bgate, const int8_t* Bup, int64_t M, int64_t N, int64_t K,
                 const GluEpilogue& glu, int dtype, int impl, cudaStream_t st);
void gemm_float(const void* A, const void* B, const void* bias, const void* residual, int act, int64_t M, int64_t N,
                int64_t K, void* C, int dtype, cudaStream_t st);

// awq.cu — AWQ-INT4 in the native (repacked, K-major) layout
struct AwqNative {
  const void* wp = nullptr;   // int32 [n, k/8]
  const void* sc = nullptr;   // f16 [n, k/group]
  const void* zr = nullptr;   // f16 [n, k/group]
  int64_t n = 0, k = 0;
  int group = 128;
  const void* sz = nullptr;   // optional: half2 {scale, zero} [k/group, n] (group-major: coalesced per-row fetches)
};
// {scale, zero} pairs of a repacked weight in group-major order (built once at load; awq_decode.cu reads it)
void awq_build_group_major(const AwqNative& w, void* sz_out /* half2 [k/group, n] */, cudaStream_t st);
// awq_decode.cu — lean decode kernel (m <= 64); false = shape not covered
bool dense_awq_decode(const void* x, const AwqNative& w, const void* bias, const void* residual, int act, int64_t m, void* y,
                      cudaStream_t st);
bool dense_awq_glu_decode(const void* x, const AwqNative& wg, const AwqNative& wu, int act, int64_t m, void* h,
                          cudaStream_t st);
// awq_gemv.cu — CUDA-core kernel for m <= 4 (no tensor cores, no barriers); false = not enabled / shape not covered
bool dense_awq_gemv(const void* x, const AwqNative& w, const void* bias, const void* residual, int act, int64_t m, void* y,
                    cudaStream_t st);
bool dense_awq_glu_gemv(const void* x, const AwqNative& wg, const AwqNative& wu, int act, int64_t m, void* h, cudaStream_t st);
void awq_repack(const int32_t* qweight, const void* scales, const int32_t* qzeros, int layout, int group, int64_t n,
                int64_t k, int32_t* wp, void* sc, void* zr, cudaStream_t st);
void awq_dequantize_ref_layout(const int32_t* qweight, const void* scales, const int32_t* qzeros, int layout, int group,
                               int64_t n, int64_t k, void* w_out, cudaStream_t st);
void awq_dequantize_native(const AwqNative& w, void* w_out, cudaStream_t st);
void dense_awq(const void* x, const AwqNative& w, const void* bias, const void* residual, int act, int64_t m, void* y,
               void* scratch_nk_f16, cudaStream_t st);
void dense_awq_glu(const void* x, const AwqNative& wg, const AwqNative& wu, int act, int64_t m, void* h,
                   void* scratch_nk_f16, void* scratch_mn_f16, cudaStream_t st);
void launch_mul_inplace_f16(void* a_inout, const void* b, int64_t n, cudaStream_t st);

// attention.cu
int attention_decode_splits(int64_t batch, int Hkv, int64_t max_len, int sm_count);
size_t attention_decode_workspace_bytes(int64_t batch, int H, int D, int splits);
void launch_attention_decode(const void* qkv, void* kc, void* vc, const float* sn, const float* cs,
                             const int32_t* lens, int64_t batch, int H, int Hkv, int D, int64_t max_len,
                             bool interleave, float scale, void* out, void* workspace, size_t workspace_bytes,
                             int splits, int dtype, cudaStream_t st);
void launch_rope_append(void* qkv, void* kc, void* vc, const float* sn, const float* cs, const int32_t* lengths,
                        int64_t batch, int64_t time, int64_t offset, int H, int Hkv, int D, int64_t max_len,
                        bool interleave, int dtype, cudaStream_t st);
void launch_attention_prefill_simple(const void* qkv, const void* kc, const void* vc, const int32_t* lengths,
                                     int64_t batch, int64_t time, int64_t offset, int H, int Hkv, int D,
                                     int64_t max_len, float scale, void* out, int dtype, cudaStream_t st);

// attention_mma.cu — tensor-core prefill attention (fp16/bf16, head_dim 64/128); false = shape not covered
bool launch_attention_prefill_mma(const void* qkv, const void* kc, const void* vc, int64_t batch, int64_t time,
                                  int64_t offset, int H, int Hkv, int D, int64_t max_len, float scale, void* out,
                                  int dtype, cudaStream_t st);
// attention_decode.cu — persistent work-balanced decode attention; false = shape not covered
bool launch_attention_decode_persistent(const void* qkv, void* kc, void* vc, const float* sn, const float* cs,
                                        const int32_t* lens, int64_t batch, int H, int Hkv, int D, int64_t max_len,
                                        bool interleave, float scale, void* out, float* partials, int32_t* tickets,
                                        int slots, int dtype, cudaStream_t st);
bool launch_attention_decode_mma(const void* qkv, void* kc, void* vc, const float* sn, const float* cs,
                                 const int32_t* lens, int64_t batch, int H, int Hkv, int D, int64_t max_len,
                                 bool interleave, float scale, void* out, float* partials, int32_t* tickets, int splits,
                                 int dtype, cudaStream_t st);
// picks the tensor-core kernel when it covers the shape (CT2B200_ATTN_PREFILL=simple forces the generic one)
void launch_attention_prefill(const void* qkv, const void* kc, const void* vc, const int32_t* lengths, int64_t batch,
                              int64_t time, int64_t offset, int H, int Hkv, int D, int64_t max_len, float scale,
                              void* out, int dtype, cudaStream_t st);

// decode_loop.cu
int sample_greedy_chunks(int64_t vocab);
void launch_sample_greedy(const void* logits, int64_t batch, int64_t vocab, const int32_t* gen, const int32_t* end_ids,
                          const int32_t* forced, int32_t* next_ids, int32_t* out_ids, int32_t* lens, float* part_v,
                          int32_t* part_i, int32_t* tickets, float* part_s, float* step_scores,
                          const int32_t* row_start, int32_t* attn_lens, int32_t* finished, int dtype, cudaStream_t st);
void launch_convert_to_f32(const void* x, int64_t n, float* y, int dtype, cudaStream_t st);
void launch_convert_from_f32(const float* x, int64_t n, void* y, int dtype, cudaStream_t st);
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t st);
void launch_mul_inplace(void* a_inout, const void* b, int64_t n, int dtype, cudaStream_t st);

// seq2seq.cu — encoder-decoder path (Translator): embeddings + positions, LayerNorm, head_dim-agnostic attention, beam search
void launch_embed_pos(const void* w, const float* w_scale, const int32_t* ids, int64_t rows, int64_t depth, float emb_scale,
                      const void* pos, int64_t time, const int32_t* step_ptr, bool zero_first, void* y, int dtype,
                      cudaStream_t st);
void launch_layer_norm(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t cols, float eps, void* y,
                       int8_t* q, float* scale, bool round, int dtype, cudaStream_t st);
void launch_attention_encoder(const void* qkv, const int32_t* lengths, int64_t batch, int S, int H, int D, float scale,
                              void* out, int dtype, cudaStream_t st);
void launch_attention_beam_self(const void* qkv, void* k_cache, void* v_cache, const int32_t* anc, const int32_t* step_ptr,
                                int64_t rows, int max_len, int H, int D, float scale, void* out, int dtype, cudaStream_t st);
void launch_attention_cross(const void* q, const void* kv, const int32_t* lengths, int64_t rows, int beam, int S, int H, int D,
                            float scale, void* out, int dtype, cudaStream_t st);
// device state of BeamSearch::search (decoding.cc:425-720); N = batch * beam rows, `stride` = allocated steps per row
struct BeamState {
  int batch = 0, beam = 1, vocab = 0, stride = 0, max_steps = 0, max_hyp = 0, max_candidates = 1, num_hypotheses = 1;
  int64_t vocab_ld = 0;             // row stride of the logits (>= vocab; a multiple of 8 lets the kernels use 16-byte accesses)
  int early_exit = 0, num_end = 0, min_length = 0;
  int start_step = 0;               // absolute position of the first search step (prompt positions come before)
  int include_eos = 1;              // DecodingOptions::include_eos_in_hypotheses
  int num_disable = 0, num_begin = 0;
  const int32_t* disable_ids = nullptr;     // SuppressTokens: disabled at every step
  const int32_t* disable_begin = nullptr;   // SuppressTokensBegin: disabled at the first search step
  // Whisper's ApplyTimestampRules (src/models/whisper.cc:742-860); ts_begin = 0 disables them
  int ts_begin = 0, ts_end = 0, ts_eot = 0, ts_no_timestamps = 0, ts_max_initial = 0;
  const int32_t* end_ids = nullptr;
  int32_t* step = nullptr;          // [1] current step, advanced by the update kernel
  int32_t* ticket = nullptr;        // [1]
  int32_t* num_finished = nullptr;  // [1] entries whose result is final (the host polls it)
  int32_t* finished = nullptr;      // [batch]
  int32_t* top_done = nullptr;      // [batch]
  int32_t* num_hyp = nullptr;       // [batch]
  int32_t* alive = nullptr;         // [2][N, stride] token history of the live beams (double-buffered by step parity)
  int32_t* anc = nullptr;           // [2][N, stride] cache slot holding position t of the row's history
  int32_t* next_ids = nullptr;      // [N] input ids of the next step
  int32_t* parent = nullptr;        // [N] or null: the row each next beam continues (gather index of Decoder::update_state)
  int32_t* hyp_tokens = nullptr;    // [batch, max_hyp, stride]
  int32_t* hyp_len = nullptr;       // [batch, max_hyp]
  float* hyp_score = nullptr;       // [batch, max_hyp] cumulative log-probability (not normalised)
};
void launch_beam_init(void* cum, int32_t* ids, int64_t rows, int beam, int start_id, int dtype, cudaStream_t st);
void launch_beam_logprobs(void* logits, const void* cum, const BeamState& s, int dtype, cudaStream_t st);
// beam <= 8: scores + the top 2 * beam of every ROW in one launch (row_scores T / row_ids int32 [batch * beam, 2 * beam], ids
// flattened over [beam, vocab]); launch_beam_update(per_row = true) merges the rows of an entry
void launch_beam_rows(void* logits, const void* cum, const BeamState& s, void* row_scores, int32_t* row_ids, int dtype,
                      cudaStream_t st);
// one prompt position without a search step: next ids = forced_next [rows], identity ancestry, step + 1
void launch_beam_force(const BeamState& s, const int32_t* forced_next, cudaStream_t st);
// out[r] = softmax(logits[r * row_stride : +vocab])[token]
void launch_token_prob(const void* logits, int64_t rows, int64_t vocab, int64_t row_stride, int token, float* out, int dtype,
                       cudaStream_t st);
// ops::Conv1D as im2col (+ the float Dense): cols [batch * Tout, Cin * K] T from x [batch, Cin, Tin] (channel_major) or [batch, Tin, Cin]
void launch_im2col(const void* x, bool x_is_f32, int64_t batch, int64_t Cin, int64_t Tin, int64_t Tout, int K, int stride,
                   int padding, bool channel_major, void* cols, int dtype, cudaStream_t st);
void launch_add_positions(void* x, const void* pos, int64_t rows, int64_t time, int64_t depth, int dtype, cudaStream_t st);
void launch_beam_update(const BeamState& s, const void* cand_scores, const int32_t* cand_ids, void* cum, bool per_row, int dtype,
                        cudaStream_t st);
// Decoder::update_state / replicate_state for the contiguous per-row caches of the decoder-only engine (decoder.cc:33-139):
// dst[row] = src[parent[row]] (parent == null: src[row / beam]) for positions [0, positions) of every kv head;
// caches [rows, Hkv, max_len, D] T
void launch_kv_gather(const void* src_k, const void* src_v, void* dst_k, void* dst_v, const int32_t* parent, int beam, int64_t rows,
                      int Hkv, int64_t max_len, int D, int64_t positions, int dtype, cudaStream_t st);
// float32 Dense (true fp32 FMAs): c [m,n] = a [m,k] . b [n,k]^T with bias / activation / residual
void gemm_f32(const float* A, const float* B, const float* bias, const float* residual, int act, int64_t M, int64_t N,
              int64_t K, float* C, cudaStream_t st);

}  // namespace ct2b200

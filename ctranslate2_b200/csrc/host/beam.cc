// beam.cc — see beam.h.
#include "beam.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace ct2b200 {

BeamSearchArena::BeamSearchArena() {
  end_ids.alloc(64 * sizeof(int32_t));
  counters.alloc(64);
  CT2_CUDA_CHECK(cudaMemset(counters.ptr, 0, 64));
  CT2_CUDA_CHECK(cudaDeviceSynchronize());   // legacy-stream memset vs the engine's non-blocking stream
}

BeamSearchArena::~BeamSearchArena() {
  if (host) cudaFreeHost(host);
}

bool BeamSearchArena::ensure(int64_t batch, int beam, int64_t steps, size_t es) {
  if (batch <= cap_batch && beam <= cap_beam && steps <= cap_steps) return false;
  cap_batch = std::max(cap_batch, batch);
  cap_beam = std::max(cap_beam, beam);
  cap_steps = std::max(cap_steps, steps);
  const int64_t B = cap_batch, L = cap_steps, N = B * cap_beam, maxh = max_hyp();
  cum.alloc(N * es);
  cand_scores.alloc(N * 2 * cap_beam * es);        // [B, 2 beam] (entry candidates) or [B * beam, 2 beam] (row candidates)
  cand_ids.alloc(N * 2 * cap_beam * 4);
  next_ids.alloc(N * 4);
  parent.alloc(N * 4);
  finished.alloc(B * 4);
  top_done.alloc(B * 4);
  num_hyp.alloc(B * 4);
  alive.alloc(2 * N * L * 4);
  anc.alloc(2 * N * L * 4);
  CT2_CUDA_CHECK(cudaMemset(anc.ptr, 0, anc.bytes));
  CT2_CUDA_CHECK(cudaDeviceSynchronize());   // legacy-stream memset vs the engine's non-blocking stream
  hyp_tokens.alloc(B * maxh * L * 4);
  hyp_len.alloc(B * maxh * 4);
  hyp_score.alloc(B * maxh * 4);
  const size_t need = static_cast<size_t>(B) * maxh * (L + 2) + B + 64;
  if (need > host_elems) {
    if (host) cudaFreeHost(host);
    CT2_CUDA_CHECK(cudaMallocHost(&host, need * sizeof(int32_t)));
    host_elems = need;
  }
  return true;
}

BeamState BeamSearchArena::state(int64_t batch, int beam, int64_t vocab, int64_t max_steps, int64_t min_length, float patience,
                                 float length_penalty, int num_hypotheses, int num_end) const {
  BeamState bs;
  bs.batch = static_cast<int>(batch);
  bs.beam = beam;
  bs.vocab = static_cast<int>(vocab);
  bs.vocab_ld = vocab;
  bs.stride = static_cast<int>(cap_steps);
  bs.max_steps = static_cast<int>(max_steps);
  bs.max_hyp = static_cast<int>(max_hyp());
  bs.min_length = static_cast<int>(min_length);
  bs.max_candidates = std::max(1, static_cast<int>(std::lround(beam * patience)));   // decoding.cc:415-418
  bs.num_hypotheses = num_hypotheses;
  bs.early_exit = length_penalty == 0.f ? 1 : 0;
  bs.num_end = num_end;
  bs.end_ids = end_ids.as<int32_t>();
  bs.step = counters.as<int32_t>();
  bs.ticket = bs.step + 1;
  bs.num_finished = bs.step + 2;
  bs.finished = finished.as<int32_t>();
  bs.top_done = top_done.as<int32_t>();
  bs.num_hyp = num_hyp.as<int32_t>();
  bs.alive = alive.as<int32_t>();
  bs.anc = anc.as<int32_t>();
  bs.next_ids = next_ids.as<int32_t>();
  bs.parent = parent.as<int32_t>();
  bs.hyp_tokens = hyp_tokens.as<int32_t>();
  bs.hyp_len = hyp_len.as<int32_t>();
  bs.hyp_score = hyp_score.as<float>();
  return bs;
}

void BeamSearchArena::reset(const BeamState& bs, int32_t start_id, int dtype, cudaStream_t st) {
  CT2_CUDA_CHECK(cudaMemsetAsync(counters.ptr, 0, 64, st));
  CT2_CUDA_CHECK(cudaMemsetAsync(finished.ptr, 0, bs.batch * 4, st));
  CT2_CUDA_CHECK(cudaMemsetAsync(top_done.ptr, 0, bs.batch * 4, st));
  CT2_CUDA_CHECK(cudaMemsetAsync(num_hyp.ptr, 0, bs.batch * 4, st));
  launch_beam_init(cum.ptr, next_ids.as<int32_t>(), static_cast<int64_t>(bs.batch) * bs.beam, bs.beam, start_id, dtype, st);
}

void BeamSearchArena::step(void* logits, const BeamState& bs, int dtype, cudaStream_t st) {
  if (bs.beam <= 8) {                              // one pass over the logits, nothing written back
    launch_beam_rows(logits, cum.ptr, bs, cand_scores.ptr, cand_ids.as<int32_t>(), dtype, st);
    launch_beam_update(bs, cand_scores.ptr, cand_ids.as<int32_t>(), cum.ptr, true, dtype, st);
    return;
  }
  CT2_REQUIRE(bs.vocab_ld == bs.vocab, "beam search with beam_size > 8 needs contiguous logits rows");
  launch_beam_logprobs(logits, cum.ptr, bs, dtype, st);
  launch_topk(logits, bs.batch, static_cast<int64_t>(bs.beam) * bs.vocab, 2 * bs.beam, cand_scores.ptr, cand_ids.as<int32_t>(), dtype,
              st);
  launch_beam_update(bs, cand_scores.ptr, cand_ids.as<int32_t>(), cum.ptr, false, dtype, st);
}

// finalize_result (decoding.cc:189-254): normalise by length^penalty, sort (stable: equal scores keep registration order), keep
// num_hypotheses, strip `strip_ids` from the tail
std::vector<TranslationHypotheses> BeamSearchArena::collect(const BeamState& bs, float length_penalty, int num_hypotheses,
                                                            const std::vector<int32_t>& strip_ids, cudaStream_t st) {
  const int64_t B = bs.batch, maxh = bs.max_hyp, stride = bs.stride;
  int32_t* h_nh = host;
  int32_t* h_len = h_nh + B;
  float* h_score = reinterpret_cast<float*>(h_len + B * maxh);
  int32_t* h_tok = h_len + 2 * B * maxh;
  CT2_CUDA_CHECK(cudaMemcpyAsync(h_nh, num_hyp.ptr, B * 4, cudaMemcpyDeviceToHost, st));
  CT2_CUDA_CHECK(cudaMemcpyAsync(h_len, hyp_len.ptr, B * maxh * 4, cudaMemcpyDeviceToHost, st));
  CT2_CUDA_CHECK(cudaMemcpyAsync(h_score, hyp_score.ptr, B * maxh * 4, cudaMemcpyDeviceToHost, st));
  CT2_CUDA_CHECK(cudaMemcpyAsync(h_tok, hyp_tokens.ptr, B * maxh * stride * 4, cudaMemcpyDeviceToHost, st));
  CT2_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<TranslationHypotheses> out(B);
  for (int64_t b = 0; b < B; ++b) {
    const int nh = h_nh[b];
    std::vector<float> sc(nh);
    for (int j = 0; j < nh; ++j) {
      const float len = static_cast<float>(h_len[b * maxh + j]);
      sc[j] = h_score[b * maxh + j] / std::pow(len, length_penalty);
    }
    std::vector<int> order(nh);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return sc[a] > sc[c]; });
    if (static_cast<int>(order.size()) > num_hypotheses) order.resize(num_hypotheses);
    for (int j : order) {
      const int32_t* t = h_tok + (b * maxh + j) * stride;
      std::vector<int32_t> toks(t, t + h_len[b * maxh + j]);
      while (!toks.empty() && std::find(strip_ids.begin(), strip_ids.end(), toks.back()) != strip_ids.end()) toks.pop_back();
      out[b].tokens.push_back(std::move(toks));
      out[b].scores.push_back(sc[j]);
    }
  }
  return out;
}

}  // namespace ct2b200

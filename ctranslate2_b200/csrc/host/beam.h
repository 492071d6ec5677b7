// beam.h — device-resident state of BeamSearch::search (src/decoding.cc:425-720) shared by the encoder-decoder engine
// (translator.cc) and the decoder-only Generator (engine.cc): buffers, the three kernels of a search step and
// finalize_result (decoding.cc:189-254) on the host.  The kernels live in kernels/seq2seq.cu.
#pragma once

#include <cstdint>
#include <vector>

#include "engine.h"

namespace ct2b200 {

struct TranslationHypotheses {            // per batch entry, best first
  std::vector<std::vector<int32_t>> tokens;
  std::vector<float> scores;
};

struct BeamSearchArena {
  int64_t cap_batch = 0, cap_steps = 0;
  int cap_beam = 0;
  DeviceBuffer cum, cand_scores, cand_ids, next_ids, end_ids, counters, finished, top_done, num_hyp, alive, anc, parent;
  DeviceBuffer hyp_tokens, hyp_len, hyp_score;
  int32_t* host = nullptr;                // pinned staging of the results
  size_t host_elems = 0;

  BeamSearchArena();
  ~BeamSearchArena();
  BeamSearchArena(const BeamSearchArena&) = delete;
  BeamSearchArena& operator=(const BeamSearchArena&) = delete;

  int64_t max_hyp() const { return 3 * static_cast<int64_t>(cap_beam); }   // round(beam * patience) + beam, patience <= 2
  // grows the buffers; true when something was reallocated (captured graphs over them are stale then)
  bool ensure(int64_t batch, int beam, int64_t steps, size_t elem_size);
  BeamState state(int64_t batch, int beam, int64_t vocab, int64_t max_steps, int64_t min_length, float patience,
                  float length_penalty, int num_hypotheses, int num_end) const;
  // clears the counters / flags and starts every beam from start_id (beam 0 live, the others at the lowest score)
  void reset(const BeamState& bs, int32_t start_id, int dtype, cudaStream_t st);
  // one search step over logits [batch * beam, vocab] T (modified in place): log-probabilities + cumulative scores,
  // TopK of 2 * beam candidates per entry, bookkeeping; next_ids / cum / parent / histories are updated on the device
  void step(void* logits, const BeamState& bs, int dtype, cudaStream_t st);
  std::vector<TranslationHypotheses> collect(const BeamState& bs, float length_penalty, int num_hypotheses,
                                             const std::vector<int32_t>& strip_ids, cudaStream_t st);
};

}  // namespace ct2b200

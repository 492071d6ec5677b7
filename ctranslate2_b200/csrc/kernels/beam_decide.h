// beam_decide.h — the per-entry bookkeeping of one BeamSearch::search step (src/decoding.cc:595-663) as a plain function shared
// by the device kernel (beam_update_kernel, seq2seq.cu: thread 0 of the entry's CTA) and a host entry point
// (ct2b200_beam_decide_host) that lets the CPU test suite drive the very same code against the oracle.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define CT2B200_HD __host__ __device__
#else
#define CT2B200_HD
#endif

namespace ct2b200 {

constexpr int kMaxBeam = 32;

struct BeamDecision {
  int active[kMaxBeam];     // candidate index (0 .. 2 * beam) each next beam continues
  int hyp_slot[kMaxBeam];   // hypothesis slot registered for candidate k, or -1
  int hyp_len[kMaxBeam];    // its length (the end token kept or dropped per include_eos_in_hypotheses)
  int num_hyp;              // hypotheses of the entry after this step
  int top_done;             // the top beam has finished (early exit without length penalty)
  int finished;             // the entry is complete
};

// word[0 .. 2 * beam): the candidates' tokens in TopK order; rel = step of the search; was_finished: results already frozen
CT2B200_HD inline void beam_decide(int beam, const int* word, const int32_t* end_ids, int num_end, int rel, int max_steps,
                                   bool was_finished, int top_done, int num_hyp, int max_hyp, int max_candidates,
                                   int num_hypotheses, int early_exit, int include_eos, BeamDecision& d) {
  const int nc = 2 * beam;
  auto is_end = [&](int w) {
    for (int e = 0; e < num_end; ++e)
      if (end_ids[e] == w) return true;
    return false;
  };
  const bool is_last = rel + 1 >= max_steps;
  int secondary = beam;
  for (int k = 0; k < beam; ++k) {
    int next = k;
    d.hyp_slot[k] = -1;
    d.hyp_len[k] = 0;
    if (!was_finished && (is_end(word[k]) || is_last)) {
      if (k == 0) top_done = 1;
      if (num_hyp < max_hyp) {
        d.hyp_slot[k] = num_hyp++;
        d.hyp_len[k] = (is_end(word[k]) && !include_eos) ? rel : rel + 1;       // decoding.cc:601-603
      }
      for (int j = secondary; j < nc; ++j)                                       // move another active beam to this position
        if (!is_end(word[j])) {
          next = j;
          secondary = j + 1;
          break;
        }
    }
    d.active[k] = next;
  }
  d.num_hyp = num_hyp;
  d.top_done = top_done;
  d.finished = was_finished ? 1 : (is_last ? 1 : early_exit ? (top_done && num_hyp >= num_hypotheses) : (num_hyp >= max_candidates));
}

}  // namespace ct2b200

"""The device-resident beam search, checked on the CPU: csrc/kernels/beam_decide.h is the per-entry bookkeeping that thread 0 of
beam_update_kernel runs (hypothesis registration, secondary candidates, early exit / patience, BeamSearch::search
decoding.cc:595-663).  The same function is exported for the host (ct2b200_beam_decide_host); here a numpy mirror of the
step's data flow (log-softmax + cumulative scores -> TopK of 2 x beam -> decide -> histories / next ids / scores -> finalize)
drives it through whole searches and must reproduce the oracle's beam_search, which tests/test_oracle.py and
test_seq2seq_fixture.py pin against the unmodified reference."""
import ctypes

import numpy as np
import pytest

from ctranslate2_b200._lib import check, lib
from oracle import ct2_oracle as O

f32 = np.float32


def decide(beam, words, end_ids, step, max_steps, max_hyp, max_cand, num_hyp, early_exit, include_eos, state):
    w = np.array(words, np.int32)
    e = np.array(end_ids, np.int32)
    st = np.array(state, np.int32)
    active, slot, hlen = np.zeros(beam, np.int32), np.zeros(beam, np.int32), np.zeros(beam, np.int32)
    p = ctypes.c_void_p
    check(lib().ct2b200_beam_decide_host(beam, w.ctypes.data_as(p), e.ctypes.data_as(p), int(e.size), step, max_steps, max_hyp,
                                         max_cand, num_hyp, int(early_exit), int(include_eos), st.ctypes.data_as(p),
                                         active.ctypes.data_as(p), slot.ctypes.data_as(p), hlen.ctypes.data_as(p)))
    return st.tolist(), active.tolist(), slot.tolist(), hlen.tolist()


def device_flow_search(step_fn, start_ids, V, beam, max_length, min_length, end_ids, length_penalty, num_hyp, patience,
                       include_eos=True):
    """numpy mirror of translator.cc / seq2seq.cu: what the kernels compute per step, with beam_decide on the host."""
    B = len(start_ids)
    N = B * beam
    lowest = np.finfo(f32).min
    ids = np.repeat(np.asarray(start_ids), beam)
    cum = np.tile(np.array([0.0] + [lowest] * (beam - 1), f32), B)
    alive = [[] for _ in range(N)]
    state = [[0, 0, 0] for _ in range(B)]                      # num_hyp, top_done, finished
    hyps = [[] for _ in range(B)]
    max_hyp = 3 * beam
    max_cand = max(1, int(np.floor(beam * patience + 0.5)))    # std::lround
    parent = np.arange(N)
    for step in range(max_length):
        logits = np.array(step_fn(ids, step, parent), f32)
        if step < min_length:
            logits[:, list(end_ids)] = lowest
        with np.errstate(over="ignore"):
            lp = (O.softmax(logits, log=True) + cum[:, None]).astype(f32).reshape(B, beam * V)
        cs, ci = O.topk(lp, 2 * beam)
        new_ids, new_cum, new_alive, parent = np.zeros(N, np.int64), np.zeros(N, f32), [None] * N, np.zeros(N, np.int64)
        for i in range(B):
            origin, word = ci[i] // V, ci[i] % V
            state[i], active, slot, hlen = decide(beam, word, end_ids, step, max_length, max_hyp, max_cand, num_hyp,
                                                  length_penalty == 0, include_eos, state[i])
            for k in range(beam):
                if slot[k] >= 0:
                    toks = alive[i * beam + origin[k]] + [int(word[k])]
                    assert slot[k] == len(hyps[i])
                    hyps[i].append((toks[:hlen[k]], float(cs[i, k])))
                c = active[k]
                row = i * beam + k
                parent[row] = i * beam + origin[c]
                new_alive[row] = alive[parent[row]] + [int(word[c])]
                new_ids[row], new_cum[row] = word[c], cs[i, c]
        ids, cum, alive = new_ids, new_cum, new_alive
        if all(s[2] for s in state):
            break
    out = []
    for i in range(B):                                         # BeamSearchArena::collect
        sc = [f32(s) / f32(f32(len(t)) ** f32(length_penalty)) for t, s in hyps[i]]
        order = sorted(range(len(sc)), key=lambda j: -sc[j])[:num_hyp]
        res = []
        for j in order:
            t = list(hyps[i][j][0])
            while t and t[-1] in end_ids:
                t.pop()
            res.append((t, float(sc[j])))
        out.append(res)
    return out


class ToyDecoder:
    """A deterministic 'decoder' whose logits depend on the row's whole history, like a real one (state gathered by parent)."""

    def __init__(self, V, seed):
        self.V, self.r = V, np.random.default_rng(seed)
        self.table = self.r.standard_normal((V, V)).astype(f32) * 2

    def reset(self, n):
        self.h = np.zeros((n, self.V), f32)

    def oracle_step(self, ids, step):
        self.h = (0.7 * self.h + self.table[ids]).astype(f32)
        return self.h + f32(0.1 * step)

    def reorder(self, index):
        self.h = self.h[index]

    def flow_step(self, ids, step, parent):
        self.h = self.h[parent]
        return self.oracle_step(ids, step)


@pytest.mark.parametrize("beam,num_hyp,lp,patience,min_len", [(1, 1, 1.0, 1.0, 0), (2, 2, 1.0, 1.0, 0), (4, 2, 0.0, 1.0, 0),
                                                              (3, 3, 0.6, 1.0, 2), (4, 4, 1.0, 2.0, 0), (5, 1, 1.0, 0.5, 3)])
def test_device_bookkeeping_reproduces_the_oracle_search(beam, num_hyp, lp, patience, min_len):
    V, B, end = 23, 3, [2, 7]
    for seed in range(6):
        dec = ToyDecoder(V, seed)
        start = np.array([1, 3, 5][:B])
        dec.reset(B * beam)
        want = O.beam_search(dec.oracle_step, dec.reorder, start, V, beam, 9, min_len, end, lp, num_hyp, patience)
        dec.reset(B * beam)
        got = device_flow_search(dec.flow_step, start, V, beam, 9, min_len, end, lp, num_hyp, patience)
        for g, w in zip(got, want):
            assert [h[0] for h in g] == [h[0] for h in w], (seed, g, w)
            np.testing.assert_allclose([h[1] for h in g], [h[1] for h in w], rtol=1e-6, atol=1e-6)


def test_hypotheses_without_the_end_token():
    """include_eos_in_hypotheses = false (Whisper, whisper.cc:309): the registered length drops the end token."""
    V, beam = 23, 3
    dec = ToyDecoder(V, 11)
    dec.reset(beam)
    want = O.beam_search(dec.oracle_step, dec.reorder, np.array([4]), V, beam, 8, 0, [2, 7], 1.0, 2, 1.0,
                         include_eos_in_hypotheses=False)
    dec.reset(beam)
    got = device_flow_search(dec.flow_step, np.array([4]), V, beam, 8, 0, [2, 7], 1.0, 2, 1.0, include_eos=False)
    assert [h[0] for h in got[0]] == [h[0] for h in want[0]]
    np.testing.assert_allclose([h[1] for h in got[0]], [h[1] for h in want[0]], rtol=1e-6)

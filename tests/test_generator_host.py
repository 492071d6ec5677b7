"""Host logic of ctranslate2_b200.Generator.generate_batch without a device: the C-ABI is replaced by a recording fake, so what
is checked is the Python side — padding of ragged prompts, re-batching of requests larger than the arena (the reference's
replica pool splits by max_batch_size, longest examples first, and answers in request order: src/batch_reader.cc), option
checks.  No compute is claimed here; the GPU tests cover the real library."""
import ctypes

import numpy as np
import pytest

import ctranslate2_b200.generator as G


def _i32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_int32 * n).from_address(ptr.value))


class FakeLib:
    """ct2b200_generate_batch: row b 'generates' max_length tokens prompt[b][0] + 1 + t, so the result identifies its prompt."""

    def __init__(self):
        self.calls = []

    def ct2b200_generate_batch(self, h, ids, lens, B, P, max_length, min_length, end_ids, n_end, return_end, out, out_lens):
        B, P, L = B.value, P.value, max_length.value
        ids_a = _i32(ids, B * P).reshape(B, P).copy()
        lens_a = _i32(lens, B).copy()
        self.calls.append((ids_a, lens_a))
        out_a, out_l = _i32(out, B * L).reshape(B, L), _i32(out_lens, B)
        for b in range(B):
            out_a[b] = ids_a[b, 0] + 1 + np.arange(L)
            out_l[b] = L
        return 0

    def ct2b200_last_error(self):
        return b""


@pytest.fixture
def gen(monkeypatch):
    fake = FakeLib()
    monkeypatch.setattr(G, "lib", lambda: fake)
    g = object.__new__(G.Generator)
    g._h, g.max_batch_size, g.max_length, g.vocab_size = 1, 4, 64, 1000
    g._tokens = ["<t%d>" % i for i in range(1000)]
    g._token_to_id, g._config = None, {"eos_token": "<t2>"}
    yield g, fake
    g._h = None


def test_requests_larger_than_the_arena_are_rebatched_longest_first_and_answered_in_order(gen):
    g, fake = gen
    r = np.random.default_rng(0)
    prompts = [[int(10 * (i + 1))] + r.integers(3, 900, size=int(n)).tolist() for i, n in enumerate(r.integers(0, 9, size=11))]
    res = g.generate_batch(prompts, max_length=3, min_length=3, end_token=[2])
    assert [x.sequences_ids[0] for x in res] == [[p[0] + 1, p[0] + 2, p[0] + 3] for p in prompts]      # request order
    assert [len(c[1]) for c in fake.calls] == [4, 4, 3]                                                # arena-sized chunks
    served = [int(n) for c in fake.calls for n in c[1]]
    assert served == sorted((len(p) for p in prompts), reverse=True)                                   # longest first
    for ids, lens in fake.calls:                                                                       # right-padded with 0
        for b in range(len(lens)):
            assert (ids[b, lens[b]:] == 0).all()
    assert res[0].sequences[0] == ["<t%d>" % t for t in res[0].sequences_ids[0]]


def test_a_request_that_fits_is_one_call(gen):
    g, fake = gen
    g.generate_batch([[5, 6], [7]], max_length=2, end_token=[2])
    assert len(fake.calls) == 1 and fake.calls[0][1].tolist() == [2, 1]
    assert g.generate_batch([], max_length=2) == [] and len(fake.calls) == 1


def test_option_checks_happen_before_any_call(gen):
    g, fake = gen
    for kw in (dict(sampling_topk=5), dict(include_prompt_in_result=True), dict(repetition_penalty=1.2), dict(beam_size=0),
               dict(no_repeat_ngram_size=1), dict(disable_unk=True), dict(not_an_option=1)):
        with pytest.raises(ValueError):
            g.generate_batch([[5]] * 9, max_length=2, **kw)
    with pytest.raises(ValueError):
        g.generate_batch([[5, 1000]], max_length=2)                 # id outside the vocabulary
    assert fake.calls == []


def test_beam_requests_keep_their_argument_errors_when_rebatched(gen):
    g, fake = gen
    with pytest.raises(ValueError):
        g.generate_batch([[1, 2, 3], [1, 2], [4, 5, 6]], max_length=4, beam_size=2)      # cap = 2 rows per call; ragged prompts
    assert fake.calls == []


def test_max_batch_size_of_the_call_splits_below_the_arena(gen):
    g, fake = gen
    prompts = [[10 * (i + 1)] * (1 + i % 3) for i in range(5)]
    res = g.generate_batch(prompts, max_length=2, min_length=2, end_token=[2], max_batch_size=2)
    assert [x.sequences_ids[0][0] for x in res] == [p[0] + 1 for p in prompts]
    assert [len(c[1]) for c in fake.calls] == [2, 2, 1]
    for bad in (-1, 1.5, True):
        with pytest.raises(ValueError):
            g.generate_batch(prompts, max_length=2, max_batch_size=bad)
    with pytest.raises(ValueError):
        g.generate_batch(prompts, max_length=2, batch_type="tokens")

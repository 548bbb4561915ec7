"""The beam-search oracle against brute force: with a beam wider than the number of prefixes the search is
exhaustive, so it must return the exact most probable labelling and its exact log probability."""
import os

import numpy as np
import pytest

from oracle import beam_oracle as bo


@pytest.mark.parametrize("T,C,seed", [(4, 3, 0), (5, 3, 1), (6, 3, 2), (5, 4, 3), (4, 4, 4), (3, 5, 5), (6, 2, 6)])
def test_exhaustive_beam_equals_brute_force(T, C, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(T, C)) * 2.0
    want_lab, want_lp, table = bo.best_labelling_brute_force(x)
    got_lab, got_lp = bo.beam_search_one(x, beam_width=10 ** 6, merge_repeated=False)
    assert got_lab == want_lab
    assert abs(got_lp - want_lp) < 1e-9
    # every labelling's probability is right, not only the best one: total mass is 1
    assert abs(np.logaddexp.reduce(list(table.values()))) < 1e-9


def test_merge_repeated_collapses_the_output_and_narrow_beams_still_decode():
    # frames strongly say: a, a, blank, a, b, b  ->  labelling [a, a, b]; TF's merge_repeated makes it [a, b]
    a, b, blank = 0, 1, 2
    x = np.full((6, 3), -4.0)
    for t, c in enumerate([a, a, blank, a, b, b]):
        x[t, c] = 4.0
    assert bo.beam_search_one(x, 100, merge_repeated=False)[0] == [a, a, b]
    assert bo.beam_search_one(x, 100, merge_repeated=True)[0] == [a, b]
    assert bo.beam_search_one(x, 1, merge_repeated=False)[0] == [a, a, b]     # greedy-width beam
    vals, offs, lps = bo.beam_search(x[:, None, :].repeat(2, 1), [6, 3], 100, False)
    assert offs.tolist() == [0, 3, 4] and vals.tolist() == [a, a, b, a] and (lps <= 0).all()


# ---- the C port (TF's trie + TopN control flow) against the dictionary oracle: two differently built programs
def _both(x, seq, W, P=1, merge=True):
    from oracle import c_oracle
    hyp, hl, lp = c_oracle.beam_search(x, seq, W, P, merge)
    for b in range(x.shape[1]):
        want = bo.beam_search_one(x[: int(seq[b]), b, :].astype(np.float64), W, merge, top_paths=P)
        for p in range(P):
            if p < len(want):
                assert hyp[b, p, : hl[b, p]].tolist() == want[p][0], (b, p)
                assert abs(lp[b, p] - want[p][1]) <= 1e-10 * max(1.0, abs(want[p][1]))
            else:
                assert hl[b, p] == 0 and lp[b, p] == -np.inf


@pytest.mark.parametrize("T,B,C,W", [(30, 3, 38, 100), (40, 3, 6, 4), (25, 2, 12, 1), (60, 2, 5, 7), (20, 2, 38, 16)])
def test_c_port_equals_dictionary_oracle(T, B, C, W):
    rng = np.random.default_rng(100 * T + W)
    x = (rng.normal(size=(T, B, C)) * 3).astype(np.float32)
    seq = np.array([T] + list(rng.integers(0, T + 1, size=B - 1)), np.int32)
    _both(x, seq, W)


def test_c_port_ties_top_paths_and_strided_input():
    _both(np.zeros((8, 2, 6), np.float32), np.array([8, 5], np.int32), W=10, P=4, merge=False)   # exact ties
    rng = np.random.default_rng(9)
    xb = (rng.normal(size=(2, 20, 7)) * 2).astype(np.float32)                                 # batch-major
    _both(xb.transpose(1, 0, 2), np.array([20, 11], np.int32), W=12, P=3, merge=True)


def test_c_port_narrow_beam_drops_and_readmits_prefixes():
    # a narrow beam on flat-ish logits keeps evicting prefixes whose parents come back later
    rng = np.random.default_rng(21)
    x = (rng.normal(size=(80, 2, 4)) * 0.7).astype(np.float32)
    _both(x, np.array([80, 64], np.int32), W=3, P=3, merge=False)
    _both(x, np.array([80, 64], np.int32), W=5, P=2, merge=True)


# ---- committed golden fixture: exact labelling probabilities of tiny utterances (tests/golden/make_beam_golden.py)
def _golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "beam_bruteforce.npz"))
    for i in range(int(g["cases"])):
        n = int(g["n%d" % i])
        want = [(g["labels%d" % i][j, : g["lens%d" % i][j]].tolist(), float(g["logp%d" % i][j])) for j in range(n)]
        yield g["x%d" % i], want


def test_both_oracles_reproduce_the_golden_fixture():
    from oracle import c_oracle
    cases = list(_golden())
    assert len(cases) >= 16 and sum(len(w) for _, w in cases) >= 40
    for x, want in cases:
        got = bo.beam_search_one(x.astype(np.float64), beam_width=10 ** 5, merge_repeated=False, top_paths=len(want))
        hyp, hl, lp = c_oracle.beam_search(x[:, None, :], [x.shape[0]], 10 ** 5, len(want), False)
        for j, (lab, v) in enumerate(want):
            assert got[j][0] == lab and abs(got[j][1] - v) < 1e-9
            assert hyp[0, j, : hl[0, j]].tolist() == lab and abs(lp[0, j] - v) < 1e-9


def _tf_published_beam_case():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "tf_ctc_decoder_ops_test_beam.json")) as f:
        g = json.load(f)
    p = np.asarray(g["input_prob_matrix"], np.float32) + np.float32(g["offset"])
    x = np.concatenate([p, np.zeros((g["padding_frames"], p.shape[1]), np.float32)])[:, None, :]
    return g, np.ascontiguousarray(x)


def test_tf_published_beam_search_case():
    """TensorFlow's own ctc_decoder_ops_test.py beam search case (tests/golden/tf_ctc_decoder_ops_test_beam.json):
    both oracles return the published hypotheses and their log probabilities to the printed digits."""
    from oracle import c_oracle
    g, x = _tf_published_beam_case()
    T = g["seq_len"]
    got = bo.beam_search_one(x[:T, 0, :], g["beam_width"], g["merge_repeated"], blank=g["blank"],
                             top_paths=g["top_paths"])
    hyp, hl, lp = c_oracle.beam_search(x, [T], g["beam_width"], g["top_paths"], g["merge_repeated"], blank=g["blank"])
    for j, (lab, v) in enumerate(zip(g["decoded"], g["log_prob"])):
        assert got[j][0] == lab and abs(got[j][1] - v) < 5e-6, (j, got[j])
        assert hyp[0, j, : hl[0, j]].tolist() == lab and abs(lp[0, j] - v) < 5e-6, (j, lp[0, j])

"""The beam-search oracle against brute force: with a beam wider than the number of prefixes the search is
exhaustive, so it must return the exact most probable labelling and its exact log probability."""
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

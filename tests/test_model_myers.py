"""The arithmetic of the lane-per-direction edit-distance kernel (tests/model_myers.py) against the oracle's plain
dynamic programme: word counts 2 / 4 / 7, truths that fill the last word, empty sides, symbols no truth holds."""
import numpy as np

from oracle import ctc_oracle as o

from model_myers import distance_one_ended, distance_two_ended


def _cases():
    rng = np.random.default_rng(12)
    out = []
    for words, mmax in ((2, 64), (4, 128), (7, 224)):
        for k in range(14):
            m = mmax if k == 0 else int(rng.integers(0, mmax + 1))
            n = int(rng.choice([0, 1, 2, 31, 32, 33, 64, int(rng.integers(0, 260))]))
            hi = int(rng.choice([2, 5, 37]))
            t = rng.integers(0, hi, m).tolist()
            if k % 3 == 0 and m and n:
                h = np.repeat(t, rng.integers(1, 3, m))[:n].tolist()
            else:
                h = rng.integers(-1, hi + 2, n).tolist()
            out.append((words, h, t))
    out.append((2, [3], [3]))
    out.append((2, [3], [4]))
    out.append((7, list(range(30)) * 3, list(range(30)) * 7))
    return out


def test_two_ended_join_matches_the_dynamic_programme():
    for words, h, t in _cases():
        want = o.levenshtein(h, t)
        assert distance_two_ended(h, t, words) == want, (words, len(h), len(t))
        assert distance_one_ended(h, t, words) == want, (words, len(h), len(t))


def test_bits_above_the_truth_do_not_reach_down():
    """A lane carries 32*WT bits whatever the truth's length: the same truth in a wider register gives the same column."""
    rng = np.random.default_rng(3)
    t = rng.integers(0, 4, 45).tolist()
    h = rng.integers(0, 5, 120).tolist()
    assert distance_two_ended(h, t, 2) == distance_two_ended(h, t, 7) == o.levenshtein(h, t)

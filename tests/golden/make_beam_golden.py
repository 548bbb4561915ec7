"""Generate the committed golden fixture for the beam search decoder: tests/golden/beam_bruteforce.npz.

Run once here (CPU container):  python tests/golden/make_beam_golden.py

TensorFlow (whose ctc_beam_search_decoder the reference calls, networks/tfnetwork.py:62) is not installable in
this image and the reference ships no vectors, so the fixture is the EXACT answer computed by an independent
method: every one of the C^T alignments of a tiny utterance is enumerated, collapsed (repeats merged, blanks
dropped) and its probability (product of torch.log_softmax rows, float64) added to its labelling.  A beam search
whose width is at least the number of possible prefixes is exhaustive, so it must return exactly the most
probable labellings and their log probabilities, in order.  This script shares no code with oracle/ or the
kernels.
"""
import itertools
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TOP = 5


def exact_labellings(x):
    T, C = x.shape
    blank = C - 1
    lp = torch.log_softmax(torch.from_numpy(x).double(), dim=1).numpy()
    table = {}
    for path in itertools.product(range(C), repeat=T):
        lab, prev = [], -1
        for c in path:
            if c != blank and c != prev:
                lab.append(c)
            prev = c
        s = float(sum(lp[t, c] for t, c in enumerate(path)))
        key = tuple(lab)
        table[key] = np.logaddexp(table[key], s) if key in table else s
    return sorted(table.items(), key=lambda kv: -kv[1])


def main():
    cases = [(3, 2), (4, 3), (5, 3), (6, 3), (7, 3), (8, 3), (4, 4), (5, 4), (3, 5), (4, 5), (6, 2), (5, 4), (7, 3),
             (4, 5), (5, 3), (8, 3)]
    out = {}
    rng = np.random.default_rng(2024)
    for i, (T, C) in enumerate(cases):
        scale = [0.5, 1.0, 2.0, 4.0][i % 4]
        x = (rng.normal(size=(T, C)) * scale).astype(np.float32)
        ranked = exact_labellings(x)
        # keep the ranking only where it is decided by more than rounding: stop at the first near-tie
        keep = []
        for j, (lab, v) in enumerate(ranked[:TOP]):
            if j + 1 < len(ranked) and abs(ranked[j + 1][1] - v) < 1e-9:
                break
            keep.append((lab, v))
        labs = np.full((TOP, T), -1, np.int64)
        lens = np.zeros(TOP, np.int32)
        lps = np.full(TOP, np.nan)
        for j, (lab, v) in enumerate(keep):
            labs[j, : len(lab)] = lab
            lens[j] = len(lab)
            lps[j] = v
        out["x%d" % i] = x
        out["labels%d" % i] = labs
        out["lens%d" % i] = lens
        out["logp%d" % i] = lps
        out["n%d" % i] = np.int32(len(keep))
    out["cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "beam_bruteforce.npz"), **out)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()

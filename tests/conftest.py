import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["small_ragged", "c38_ragged", "c38_full_peaky", "wide_vocab", "long_states"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """A fresh checkout has no built artefacts (they are git-ignored): compile libnasr_ctc.so for sm_100a once if it
    is missing (nvcc cross-compiles without a GPU).  Never a fallback: the tests still go through the library."""
    from neuralasr_b200 import _build
    if not os.path.exists(_build.LIB_PATH):
        _build.build_library(force=True)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


def make_batch(seed, T, B, C, Lmax, mode="ragged", scale=3.0, peaky=False, Lmin=None,
               repeat_p=0.1, empty_row=True):
    """Seeded synthetic batch as SURVEY.md §8(d) describes.  mode: full | ragged | tight."""
    rng = np.random.default_rng(seed)
    blank = C - 1
    Lmin = max(1, Lmax // 2) if Lmin is None else Lmin
    L = rng.integers(Lmin, Lmax + 1, size=B)
    L[0] = Lmax
    if empty_row and B > 2:
        L[1] = 0
    labs = []
    for l in L:
        lab = rng.integers(0, blank, size=l)
        for i in range(1, l):
            if rng.random() < repeat_p:
                lab[i] = lab[i - 1]
        labs.append(lab.astype(np.int32))
    need = np.array([len(l) + int(np.count_nonzero(l[1:] == l[:-1])) for l in labs])
    assert need.max() <= T, "T too small for the labels"
    if mode == "full":
        seq = np.full(B, T, dtype=np.int32)
    elif mode == "tight":
        seq = np.maximum(need, 1).astype(np.int32)
    else:
        seq = np.array([rng.integers(max(n, 1, T // 2) if max(n, 1) <= T // 2 else max(n, 1), T + 1)
                        for n in need], dtype=np.int32)
        seq[0] = T
    x = (rng.standard_normal(size=(T, B, C), dtype=np.float32) * np.float32(scale))
    if peaky:
        for b in range(B):
            t = 0
            for lab in labs[b]:
                for _ in range(int(rng.integers(1, 4))):
                    if t < seq[b]:
                        x[t, b, lab] += 8.0
                        t += 1
                if t < seq[b]:
                    x[t, b, blank] += 8.0
                    t += 1
            while t < seq[b]:
                x[t, b, blank] += 8.0
                t += 1
    vals = np.concatenate(labs).astype(np.int32)
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(L)
    return dict(logits=x, label_values=vals, label_offsets=offs, seq_len=seq,
                labels_dense=labs, T=T, B=B, C=C)

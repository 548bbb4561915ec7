"""The throughput kernel's arithmetic, restated in numpy (tests/model_bfp.py), against the oracle: ratio
units, (blank,label) pairs, mirrored backward recursion, block floating point with per-lane exponents,
meet in the middle with recomputed rows kept as high words.  Runs on the CPU; it validates the
reformulation, the GPU tests validate the kernel."""
import numpy as np
import pytest

from model_bfp import ctc_loss_grad_model
from oracle import ctc_oracle


@pytest.mark.parametrize("T,L,C,NL,split", [
    (60, 10, 38, 2, None),      # forward only, backward walks the whole utterance in phase 2
    (60, 10, 38, 2, 32),
    (200, 40, 38, 4, 96),
    (90, 0, 7, 2, 48),          # empty transcript
    (400, 100, 38, 4, 208),     # long labels: wide dynamic range inside a lane
])
def test_model_matches_oracle(T, L, C, NL, split):
    rng = np.random.default_rng(T + L)
    x = (rng.standard_normal((T, C)) * 3).astype(np.float32)
    lab = rng.integers(0, C - 1, size=L)
    if L > 3:
        lab[3] = lab[2]
    loss, grad, info = ctc_loss_grad_model(x, lab, C - 1, NL=NL, K=16, split=split)
    want_loss, want_grad = ctc_oracle.ctc_loss_grad_one_vec(x.astype(np.float64), lab, C - 1)[:2]
    assert not info["alarm"]
    assert abs(loss - want_loss) <= 1e-6 * abs(want_loss)
    assert np.abs(grad - want_grad).max() < 1e-5

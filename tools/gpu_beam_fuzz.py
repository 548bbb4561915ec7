"""Randomised comparison of the beam search kernel with the C port over many small shapes (seeded)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from neuralasr_b200.networks import common  # noqa: E402
from oracle import c_oracle  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(20240)
bad = 0
for case in range(n_cases):
    C = int(rng.choice([2, 3, 5, 12, 38, 41, 64, 65, 100, 300, 1024, 1600, 4100, 8192]))
    T = int(rng.integers(1, 70))
    B = int(rng.integers(1, 5))
    W = int(rng.choice([1, 2, 3, 7, 16, 100, 128, 300]))
    P = int(min(W, rng.integers(1, 4)))
    merge = bool(rng.integers(0, 2))
    kind = case % 4
    if kind == 0:
        x = (rng.normal(size=(T, B, C)) * rng.choice([0.3, 1.0, 3.0, 8.0])).astype(np.float32)
    elif kind == 1:                                   # planted alignment
        x = rng.normal(size=(T, B, C)).astype(np.float32)
        cls = rng.integers(0, C, size=(T, B))
        np.put_along_axis(x, cls[:, :, None], 8.0, axis=2)
    elif kind == 2:                                   # quantised logits: many exact ties
        x = rng.integers(-2, 3, size=(T, B, C)).astype(np.float32)
    else:                                             # extreme range
        x = (rng.normal(size=(T, B, C)) * 40).astype(np.float32)
    seq = rng.integers(0, T + 1, size=B).astype(np.int32)
    seq[0] = T
    blank = C - 1 if case % 5 else int(rng.integers(0, C))
    try:
        dec, lp = common.beam_decoding(torch.from_numpy(x).cuda(), seq, beam_width=W, top_paths=P,
                                       merge_repeated=merge, blank=blank)
    except Exception as e:
        print("case %d unsupported: %s" % (case, str(e)[:120]))
        continue
    hyp, hl, want, margin = c_oracle.beam_search(x, seq, W, P, merge, blank=blank, with_margin=True)
    lp = lp.cpu().numpy()
    for p in range(P):
        gh, gl = dec[p].hyp.cpu().numpy(), dec[p].hyp_len.cpu().numpy()
        for b in range(B):
            if margin[b, 0] < 1e-9 or (kind == 2 and margin[b, 1] > 0):
                continue                      # a decision within rounding: may legitimately differ
            ok = gl[b] == hl[b, p] and np.array_equal(gh[b, : gl[b]], hyp[b, p, : hl[b, p]])
            if np.isfinite(want[b, p]):
                ok = ok and abs(lp[b, p] - want[b, p]) <= 1e-6 * max(1.0, abs(want[b, p]))
            else:
                ok = ok and lp[b, p] == want[b, p]
            if not ok:
                bad += 1
                print("MISMATCH case %d (kind %d) T=%d B=%d C=%d W=%d P=%d merge=%d blank=%d: b=%d p=%d  %s vs %s  %r vs %r"
                      % (case, kind, T, B, C, W, P, merge, blank, b, p, gh[b, : gl[b]].tolist(),
                         hyp[b, p, : hl[b, p]].tolist(), lp[b, p], want[b, p]), flush=True)
print("%d cases, %d mismatching paths" % (n_cases, bad))

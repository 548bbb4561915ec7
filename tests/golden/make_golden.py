"""Generate the committed golden fixtures for the CTC path.

Run once here (CPU container):  python tests/golden/make_golden.py

The reference (zeahmed/NeuralASR) delegates this path to TensorFlow 1.x, which is not
installable in this image, and ships no golden vectors of its own.  The fixtures are
therefore produced by an implementation that is independent of both our oracle and our
CUDA kernels:
  * loss / d(loss)/d(logits): torch.nn.functional.ctc_loss on CPU in float64 with
    blank = C-1 and autograd through log_softmax (== tf.nn.ctc_loss + _CTCLossGrad with
    unit upstream gradient, for feasible utterances);
  * greedy decode: torch.argmax + unique_consecutive (no ties in these inputs);
  * edit distance: torchaudio.functional.edit_distance.
Inputs are stored as float32 (what the kernels consume); expectations as float64.
"""
import os

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(seed, T, B, C, Lmax, ragged=True, scale=3.0, peaky=False):
    rng = np.random.default_rng(seed)
    blank = C - 1
    L = rng.integers(max(1, Lmax // 2), Lmax + 1, size=B)
    L[0] = Lmax
    if B > 2:
        L[1] = 0            # empty transcript is legal (l' = [blank])
    labs = []
    for l in L:
        lab = rng.integers(0, blank, size=l)
        for i in range(1, l):   # repeats exercise the "no skip over equal labels" rule
            if rng.random() < 0.15:
                lab[i] = lab[i - 1]
        labs.append(lab.astype(np.int32))
    need = np.array([len(l) + int(np.count_nonzero(l[1:] == l[:-1])) for l in labs])
    assert need.max() <= T
    if ragged:
        seq = np.array([rng.integers(max(n, 1), T + 1) for n in need], dtype=np.int32)
        seq[0] = T
        if B > 3:
            seq[3] = max(need[3], 1)   # tight: exactly the minimum feasible length
    else:
        seq = np.full(B, T, dtype=np.int32)
    x = (rng.normal(size=(T, B, C)) * scale).astype(np.float32)
    if peaky:
        for b in range(B):          # planted alignment: label runs separated by blanks
            t = 0
            for lab in labs[b]:
                for _ in range(int(rng.integers(1, 3))):
                    if t < seq[b]:
                        x[t, b, lab] += 8.0
                        t += 1
                if t < seq[b]:
                    x[t, b, blank] += 8.0
                    t += 1
            while t < seq[b]:
                x[t, b, blank] += 8.0
                t += 1
    vals = np.concatenate(labs).astype(np.int32) if len(labs) else np.zeros(0, np.int32)
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(L)

    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(xt, dim=-1)
    loss = torch.nn.functional.ctc_loss(
        lp, torch.tensor(vals, dtype=torch.long), torch.tensor(seq, dtype=torch.long),
        torch.tensor(L, dtype=torch.long), blank=blank, reduction="none", zero_infinity=False)
    loss.sum().backward()
    grad = xt.grad.numpy().copy()
    for b in range(B):              # torch leaves softmax-free zeros there already; make it explicit
        grad[seq[b]:, b, :] = 0.0

    hyp_vals, hyp_offs, nsl = [], [0], []
    for b in range(B):
        row = torch.tensor(x[:seq[b], b, :])
        am = torch.argmax(row, dim=1)
        nsl.append(float(-(row.max(dim=1).values.double().sum())))
        col = torch.unique_consecutive(am)
        col = col[col != blank].numpy().astype(np.int64)
        hyp_vals.append(col)
        hyp_offs.append(hyp_offs[-1] + len(col))
    dist = np.array([torchaudio.functional.edit_distance(list(labs[b]), list(hyp_vals[b]))
                     for b in range(B)], dtype=np.int32)
    return dict(logits=x, label_values=vals, label_offsets=offs, seq_len=seq,
                loss=loss.detach().numpy(), grad=grad,
                hyp_values=np.concatenate(hyp_vals).astype(np.int64),
                hyp_offsets=np.asarray(hyp_offs, dtype=np.int32),
                neg_sum_logits=np.asarray(nsl, dtype=np.float64), dist=dist)


CASES = {
    "small_ragged": dict(seed=1, T=30, B=5, C=7, Lmax=8),
    "c38_ragged": dict(seed=2, T=96, B=6, C=38, Lmax=24),
    "c38_full_peaky": dict(seed=3, T=120, B=4, C=38, Lmax=30, ragged=False, peaky=True),
    "wide_vocab": dict(seed=4, T=40, B=3, C=200, Lmax=12),
    "long_states": dict(seed=5, T=300, B=2, C=12, Lmax=140, scale=2.0),
}

if __name__ == "__main__":
    for name, kw in CASES.items():
        case = make_case(**kw)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **case)
        print(name, os.path.getsize(path), "bytes", "loss[0]=%.6f" % case["loss"][0])

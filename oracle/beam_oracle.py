"""CPU oracle of TensorFlow's CTC beam search decoder — TEST INFRASTRUCTURE, never shipped code.

Restates ``tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width=100, top_paths=1,
merge_repeated=True)``, the decoder the reference's ``create_model`` actually runs
(``networks/tfnetwork.py:62``), after TF 1.x ``core/util/ctc/ctc_beam_search.h`` (``CTCBeamSearchDecoder::Step``
and ``TopPaths``) as summarised in SURVEY.md A.2.  TensorFlow is not available here, so **parity with TF itself
is unpinned**; what pins this file is ``tests/test_beam_oracle.py``: with a beam wider than the number of
prefixes the search is exhaustive and must return the exact most probable labelling and its exact log
probability, which a brute-force enumeration of all alignments provides for tiny cases.

Semantics restated:
  * per frame the logits row is turned into log-softmax (newer TF; older TF only subtracts the row maximum,
    which shifts every candidate of a frame alike: same beams, same paths, a different returned score);
  * a beam entry is a prefix (tree node) with log P(prefix ends in blank), log P(prefix ends in its last label),
    and their log-sum; blank is the last class;
  * step: every active entry b keeps its prefix:  label' = LSE(label + lp[last(b)], parent term + lp[last(b)]),
    the second only if b's parent prefix is active,  parent term = parent.blank if last(b) == last(parent) else
    parent.total,  blank' = total + lp[blank],  total' = log(e^blank' + e^(label + lp) + e^(parent term + lp))
    (one three-way log-sum-exp: the same value as TF's nested LogSumExp up to rounding, and a shorter
    dependency chain in the kernel);  every extension (b, l) whose prefix is not already active enters with
    label' = lp[l] + (b.blank if l == last(b) else b.total), blank' = -inf;  the best ``beam_width`` entries by
    total survive (TF's incremental TopN with its candidate test yields exactly this set, ties aside);
  * result: the best entry's label sequence, consecutive repeats collapsed when ``merge_repeated`` (TF's
    quirk: it merges in the OUTPUT), and its log probability (TF returns ``newp.total``).
Arithmetic is float64 (TF: float32); the CUDA kernel uses float64 with the same formulas.
"""
from __future__ import annotations

import numpy as np

NEG = -np.inf
_M64 = (1 << 64) - 1
ROOT_HASH = 0x243F6A8885A308D3


def mix64(x):
    x &= _M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & _M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & _M64
    x ^= x >> 31
    return x


def child_hash(h, label):
    return mix64(h + 0x9E3779B97F4A7C15 * (label + 1))


def prefix_hash(prefix):
    h = ROOT_HASH
    for l in prefix:
        h = child_hash(h, int(l))
    return h


def _lse(a, b):
    if a == NEG:
        return b
    if b == NEG:
        return a
    m, n = (a, b) if a > b else (b, a)
    return m + np.log1p(np.exp(n - m))


def log_softmax_row(x):
    x = np.asarray(x, dtype=np.float64)
    m = x.max()
    return x - m - np.log(np.exp(x - m).sum())


def beam_search_one(x, beam_width=100, merge_repeated=True, blank=None, top_paths=None):
    """x [Tb, C] logits of one utterance -> (labels list, log_prob); with ``top_paths`` a list of such pairs."""
    Tb, C = x.shape
    blank = C - 1 if blank is None else blank
    # entry: dict(prefix tuple) -> (blank, label, total); parent = prefix[:-1]
    beams = {(): (0.0, NEG, 0.0)}
    hashes = {(): ROOT_HASH}
    for t in range(Tb):
        lp = log_softmax_row(x[t])
        cand = {}
        for pre, (pb, pl, pt) in beams.items():
            b1 = b2 = NEG
            if pre:
                par = beams.get(pre[:-1])
                b1 = pl + lp[pre[-1]]
                if par is not None:
                    b2 = (par[0] if (len(pre) >= 2 and pre[-1] == pre[-2]) else par[2]) + lp[pre[-1]]
            nb = pt + lp[blank]
            nl = _lse(b1, b2)
            m = max(nb, b1, b2)
            tot = NEG if m == NEG else m + np.log(np.exp(nb - m) + np.exp(b1 - m) + np.exp(b2 - m))
            if tot > NEG:
                cand[pre] = (nb, nl, tot, 0, hashes[pre])
        for pre, (pb, pl, pt) in beams.items():
            last = pre[-1] if pre else -1
            for l in range(C):
                if l == blank:
                    continue
                child = pre + (l,)
                if child in beams:          # already active: handled above, parent term included
                    continue
                prev = pb if l == last else pt
                v = lp[l] + prev
                if v > NEG:
                    cand[child] = (NEG, v, v, 1, child_hash(hashes[pre], l))
        # higher total first; ties: kept prefix before new extension, then the smaller prefix hash
        keep = sorted(cand.items(), key=lambda kv: (-kv[1][2], kv[1][3], kv[1][4]))[:beam_width]
        beams = {k: v[:3] for k, v in keep}
        hashes = {k: v[4] for k, v in keep}
    order = sorted(beams.items(), key=lambda kv: (-kv[1][2], hashes[kv[0]]))
    out = []
    for best, (bb, bl, bt) in order[: (top_paths or 1)]:
        labels = list(best)
        if merge_repeated:
            labels = [l for i, l in enumerate(labels) if i == 0 or l != labels[i - 1]]
        out.append((labels, float(bt)))
    return out if top_paths else out[0]


def beam_search(logits, seq_len, beam_width=100, merge_repeated=True, blank=None):
    """logits [T,B,C], seq_len [B] -> (values int64[M], offsets int32[B+1], log_prob float64[B])."""
    T, B, C = logits.shape
    vals, offs, lps = [], [0], []
    for b in range(B):
        lab, lp = beam_search_one(np.asarray(logits[: int(seq_len[b]), b, :]), beam_width, merge_repeated, blank)
        vals.extend(lab)
        offs.append(len(vals))
        lps.append(lp)
    return np.asarray(vals, np.int64), np.asarray(offs, np.int32), np.asarray(lps, np.float64)


def best_labelling_brute_force(x, blank=None):
    """Exact most probable labelling of a tiny utterance by summing ALL alignments (C**T of them)."""
    import itertools
    Tb, C = x.shape
    blank = C - 1 if blank is None else blank
    lp = np.stack([log_softmax_row(r) for r in x])
    tot = {}
    for path in itertools.product(range(C), repeat=Tb):
        lab, prev = [], -1
        for c in path:
            if c != blank and c != prev:
                lab.append(c)
            prev = c
        s = sum(lp[t, c] for t, c in enumerate(path))
        key = tuple(lab)
        tot[key] = np.logaddexp(tot.get(key, NEG), s)
    best = max(tot.items(), key=lambda kv: kv[1])
    return list(best[0]), float(best[1]), tot

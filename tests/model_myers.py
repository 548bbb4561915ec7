"""CPU model of the arithmetic of edit_distance_lanes_kernel (csrc/ctc_decode.cu): Myers' bit-vector column as one
WT-word integer per (utterance, direction), no running score, the two directions joined from the vertical deltas of
their last columns.  Python integers stand for the WT 32-bit registers of a lane (masked to 32*WT bits, so the bits above
the truth behave as in the kernel).  Test infrastructure: checked against oracle.ctc_oracle.levenshtein by
tests/test_model_myers.py; the kernel itself is checked against the C oracle on the GPU."""


def last_column(hyp, truth, words):
    """Vertical deltas (Pv, Mv) of the last column of the DP lattice of `hyp` (text) against `truth` (pattern):
    D[len(hyp)][j+1] - D[len(hyp)][j] = bit j of Pv - bit j of Mv.  One step per hypothesis symbol, as in the kernel:
    Xv = Eq | Mv; Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq; Ph = Mv | ~(Xh | Pv); Mh = Pv & Xh; shift both one row down
    (the horizontal delta entering row 0 is +1); Pv = Mh | ~(Xv | Ph); Mv = Ph & Xv."""
    full = (1 << (32 * words)) - 1
    assert len(truth) <= 32 * words
    peq = {}
    for j, v in enumerate(truth):
        peq[v] = peq.get(v, 0) | (1 << j)
    pv, mv = full, 0
    for c in hyp:
        eq = peq.get(c, 0)                       # a symbol no truth holds: the zero row of the table
        xv = eq | mv
        xh = ((((eq & pv) + pv) & full) ^ pv) | eq   # the carry out of the top word is dropped (addc without .cc)
        ph = mv | (~(xh | pv) & full)
        mh = pv & xh
        ph = ((ph << 1) | 1) & full
        mh = (mh << 1) & full
        pv = mh | (~(xv | ph) & full)
        mv = ph & xv
    return pv, mv


def below(vec, x):
    """Sum of the vertical deltas at positions < x."""
    pv, mv = vec
    mask = (1 << x) - 1
    return bin(pv & mask).count("1") - bin(mv & mask).count("1")


def distance_two_ended(hyp, truth, words):
    """The kernel's result for one utterance: forward half of the hypothesis against the truth, reversed second half
    against the reversed truth, S[j] = n + (forward deltas below j) + (backward deltas below m - j), minimum over j.
    The two lanes of an utterance split the range of j at a multiple of 32; the model walks it like they do."""
    n, m = len(hyp), len(truth)
    if n == 0 or m == 0:
        return n + m
    n1 = n - n // 2
    fwd = last_column(hyp[:n1], truth, words)
    bwd = last_column(hyp[n1:][::-1], truth[::-1], words)
    mh = (m // 2) & ~31
    best = None
    for j0, j1 in ((0, mh), (mh, m)):
        s = n + below(fwd, j0) + below(bwd, m - j0)
        best = s if best is None else min(best, s)
        for j in range(j0, j1):
            q = m - 1 - j
            s += ((fwd[0] >> j) & 1) - ((fwd[1] >> j) & 1) - ((bwd[0] >> q) & 1) + ((bwd[1] >> q) & 1)
            best = min(best, s)
    return best


def distance_one_ended(hyp, truth, words):
    """D[n][m] = n + popc(Pv) - popc(Mv) over the truth's bits of the last column."""
    if not hyp or not truth:
        return len(hyp) + len(truth)
    return len(hyp) + below(last_column(hyp, truth, words), len(truth))

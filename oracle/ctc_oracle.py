"""CPU oracle (numpy) for the CTC loss / greedy decode / label-error-rate path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``neuralasr_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may.  The product path is the CUDA library and it
fails loudly when that library is missing.

What it restates
----------------
The reference (zeahmed/NeuralASR) has no arithmetic of its own on this path:
``networks/tfnetwork.py:58-59`` (``create_loss``) calls ``tf.nn.ctc_loss``,
``networks/tfnetwork.py:61-64`` (``create_model``) calls the CTC decoder
(greedy variant is the commented alternative on line 63) and
``networks/tfnetwork.py:66-70`` (``create_metric``) calls ``tf.edit_distance``.
The algorithm therefore lives in a third-party dependency that is absent from
``/root/reference``: **TensorFlow 1.x, version unpinned** (``requirements.txt:1``
does not even list it).  This file restates TensorFlow's published CPU
algorithm (``core/util/ctc/ctc_loss_calculator.{h,cc}``, ``ctc_decoder.h``,
``core/kernels/edit_distance_op.cc``, ``python/ops/ctc_ops.py``) as summarised
in SURVEY.md Appendix A, with the defaults the reference's call sites use:
time-major inputs, blank = C-1, ``ctc_merge_repeated=True``,
``preprocess_collapse_repeated=False``, ``merge_repeated=True`` (decoder),
``normalize=True`` (edit distance).

Parity pin
----------
**Parity unpinned by the reference**: the reference ships no tests, golden
vectors or fixtures for this path, and TensorFlow cannot be imported in this
image.  The oracle is instead pinned (``tests/test_oracle.py``) by
 (i) an independent implementation: ``torch.nn.functional.ctc_loss`` on CPU in
     float64 (+ autograd through ``log_softmax``) for loss and logit gradient,
     and ``torchaudio.functional.edit_distance`` for Levenshtein, captured as
     committed fixtures under ``tests/golden/`` (``make_golden.py``);
 (ii) closed-form known answers (T=1; uniform logits path counting; greedy
     collapse; empty-side LER conventions).
Items of TF behaviour that could not be re-checked without a TF install
(T_b == 0 handling, gradient of an infeasible utterance, empty-side LER values)
are this oracle's *definition*; they are marked ``(definition)`` below.
"""
from __future__ import annotations

import numpy as np

# Per-utterance status bit flags, shared with include/nasr_ctc.h.
STATUS_OK = 0
STATUS_LABEL_OUT_OF_RANGE = 1   # label < 0, label >= C or label == blank
STATUS_SEQ_LEN_OUT_OF_RANGE = 2  # seq_len < 0 or seq_len > T
STATUS_NOT_ENOUGH_TIME = 4      # seq_len < L + repeats
STATUS_NO_VALID_PATH = 8        # log p == -inf (underflow / bypassed checks)

NEG_INF = -np.inf


def sparse_to_csr(labels, batch_size=None):
    """``(indices i64[N,2], values i32[N], shape i64[2])`` -> ``(values, offsets[B+1])``.

    Follows the layout ``utils.py:44-58`` (``sparse_tuple_from``) produces:
    batch-ordered, row-major, so offsets are a prefix sum of per-row counts.
    """
    indices, values, shape = labels
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    values = np.asarray(values, dtype=np.int32).reshape(-1)
    B = int(shape[0]) if batch_size is None else int(batch_size)
    if indices.shape[0] != values.shape[0]:
        raise ValueError("labels: indices and values disagree on N")
    rows = indices[:, 0]
    if rows.size and (np.any(rows < 0) or np.any(rows >= B)):
        raise ValueError("labels: batch index out of range")
    if rows.size > 1 and np.any(np.diff(rows) < 0):
        raise ValueError("labels: indices are not ordered by batch")
    counts = np.bincount(rows, minlength=B).astype(np.int64)
    offsets = np.zeros(B + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    return values, offsets


def _lse2(a, b):
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    m, n = (a, b) if a > b else (b, a)
    return m + np.log1p(np.exp(n - m))


def required_time(lab):
    """L + number of adjacent repeats (TF "Not enough time for target transition sequence")."""
    lab = np.asarray(lab)
    return int(lab.size + np.count_nonzero(lab[1:] == lab[:-1]))


def ctc_loss_grad_one(x, lab, blank, dtype=np.float64):
    """One utterance.  ``x``: [T_b, C] logits (already cut to seq_len); ``lab``: [L].

    Returns ``(loss, grad[T_b, C], status)``.  Scalar-loop restatement of
    ``CTCLossCalculator::CalculateLoss`` for one batch element
     — used for small cases and as the definition.
    """
    x = np.asarray(x, dtype=dtype)
    Tb, C = x.shape
    lab = np.asarray(lab, dtype=np.int64)
    L = lab.size
    status = STATUS_OK
    if L and (lab.min() < 0 or lab.max() >= C or (lab == blank).any()):
        return np.inf, np.zeros_like(x), STATUS_LABEL_OUT_OF_RANGE
    if Tb == 0:
        # (definition) zero-length utterance: skipped, loss 0, gradient 0.
        return 0.0, np.zeros_like(x), status
    if Tb < required_time(lab):
        status |= STATUS_NOT_ENOUGH_TIME
    U = 2 * L + 1
    lp = np.full(U, blank, dtype=np.int64)
    lp[1::2] = lab
    # softmax, then log of the normalised probability (TF takes log(y)).
    xm = x - x.max(axis=1, keepdims=True)
    e = np.exp(xm)
    y = e / e.sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore"):
        logy = np.log(y)
    alpha = np.full((Tb, U), NEG_INF, dtype=dtype)
    beta = np.full((Tb, U), NEG_INF, dtype=dtype)
    alpha[0, 0] = logy[0, blank]
    if U > 1:
        alpha[0, 1] = logy[0, lp[1]]
    for t in range(1, Tb):
        lo = max(0, U - 2 * (Tb - t))
        hi = min(U, 2 * (t + 1))
        for u in range(lo, hi):
            s = alpha[t - 1, u]
            if u > 0:
                s = _lse2(s, alpha[t - 1, u - 1])
            if u > 1 and lp[u] != blank and lp[u] != lp[u - 2]:
                s = _lse2(s, alpha[t - 1, u - 2])
            alpha[t, u] = logy[t, lp[u]] + s
    for u in range(max(0, U - 2), U):
        beta[Tb - 1, u] = 0.0
    for t in range(Tb - 2, -1, -1):
        lo = max(0, U - 2 * (Tb - t))
        hi = min(U, 2 * (t + 1))
        for u in range(lo, hi):
            s = beta[t + 1, u] + logy[t + 1, lp[u]]
            if u + 1 < U:
                s = _lse2(s, beta[t + 1, u + 1] + logy[t + 1, lp[u + 1]])
            if u + 2 < U and lp[u] != blank and lp[u] != lp[u + 2]:
                s = _lse2(s, beta[t + 1, u + 2] + logy[t + 1, lp[u + 2]])
            beta[t, u] = s
    logp = NEG_INF
    for u in range(U):
        logp = _lse2(logp, alpha[0, u] + beta[0, u])
    if logp == NEG_INF:
        # (definition) no valid path: loss = +inf, gradient = softmax.
        return np.inf, y.copy(), status | STATUS_NO_VALID_PATH
    grad = y.copy()
    for t in range(Tb):
        occ = np.full(C, NEG_INF, dtype=dtype)
        for u in range(U):
            occ[lp[u]] = _lse2(occ[lp[u]], alpha[t, u] + beta[t, u])
        grad[t] -= np.exp(occ - logp)
    return float(-logp), grad, status


def _logaddexp_cols(a, b):
    return np.logaddexp(a, b)


def ctc_loss_grad_one_vec(x, lab, blank, dtype=np.float64):
    """Vectorised-over-states version of :func:`ctc_loss_grad_one` (same maths,
    whole state vector per time step; band limits omitted because they only skip
    states that are provably -inf).  Used for the larger parity cases.
    """
    x = np.asarray(x, dtype=dtype)
    Tb, C = x.shape
    lab = np.asarray(lab, dtype=np.int64)
    L = lab.size
    status = STATUS_OK
    if L and (lab.min() < 0 or lab.max() >= C or (lab == blank).any()):
        return np.inf, np.zeros_like(x), STATUS_LABEL_OUT_OF_RANGE
    if Tb == 0:
        return 0.0, np.zeros_like(x), status
    if Tb < required_time(lab):
        status |= STATUS_NOT_ENOUGH_TIME
    U = 2 * L + 1
    lp = np.full(U, blank, dtype=np.int64)
    lp[1::2] = lab
    skip = np.zeros(U, dtype=bool)           # may take the u-2 transition
    skip[2:] = (lp[2:] != blank) & (lp[2:] != lp[:-2])
    xm = x - x.max(axis=1, keepdims=True)
    e = np.exp(xm)
    y = e / e.sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore"):
        logy = np.log(y)
    em = logy[:, lp]                          # [Tb, U]
    alpha = np.full((Tb, U), NEG_INF, dtype=dtype)
    beta = np.full((Tb, U), NEG_INF, dtype=dtype)
    alpha[0, 0] = em[0, 0]
    if U > 1:
        alpha[0, 1] = em[0, 1]
    ninf1 = np.full(1, NEG_INF, dtype=dtype)
    ninf2 = np.full(2, NEG_INF, dtype=dtype)
    with np.errstate(invalid="ignore"):
        for t in range(1, Tb):
            a = alpha[t - 1]
            s = np.logaddexp(a, np.concatenate([ninf1, a[:-1]]))
            a2 = np.concatenate([ninf2, a[:-2]])[:U]
            s = np.where(skip, np.logaddexp(s, a2), s)
            alpha[t] = em[t] + s
        beta[Tb - 1, max(0, U - 2):] = 0.0
        skip_fwd = np.zeros(U, dtype=bool)    # u may jump to u+2
        skip_fwd[:-2] = skip[2:]
        for t in range(Tb - 2, -1, -1):
            b = beta[t + 1] + em[t + 1]
            s = np.logaddexp(b, np.concatenate([b[1:], ninf1]))
            b2 = np.concatenate([b[2:], ninf2])[:U]
            s = np.where(skip_fwd, np.logaddexp(s, b2), s)
            beta[t] = s
        ab0 = alpha[0] + beta[0]
    ab0 = ab0[np.isfinite(ab0)]
    if ab0.size == 0:
        return np.inf, y.copy(), status | STATUS_NO_VALID_PATH
    m = ab0.max()
    logp = m + np.log(np.exp(ab0 - m).sum())
    with np.errstate(invalid="ignore"):
        post = np.exp(alpha + beta - logp)    # [Tb, U]; -inf + x -> 0
    post = np.nan_to_num(post, nan=0.0)
    occ = np.zeros((Tb, C), dtype=dtype)
    np.add.at(occ, (slice(None), lp), post)
    return float(-logp), y - occ, status


def ctc_loss_grad(logits, label_values, label_offsets, seq_len, blank=None,
                  dtype=np.float64, vectorised=True):
    """Batch CTC loss and logit gradient — ``tf.nn.ctc_loss`` + ``_CTCLossGrad``
    with unit upstream gradient (call site
    ``networks/tfnetwork.py:59``).

    ``logits`` [T,B,C] time-major; labels in CSR form; ``seq_len`` [B].
    Returns ``(loss[B], grad[T,B,C], status[B])``.  Frames ``t >= seq_len[b]``
    get exactly zero gradient.
    """
    logits = np.asarray(logits)
    T, B, C = logits.shape
    if blank is None:
        blank = C - 1
    fn = ctc_loss_grad_one_vec if vectorised else ctc_loss_grad_one
    loss = np.zeros(B, dtype=dtype)
    grad = np.zeros((T, B, C), dtype=dtype)
    status = np.zeros(B, dtype=np.int32)
    for b in range(B):
        Tb = int(seq_len[b])
        lab = label_values[label_offsets[b]:label_offsets[b + 1]]
        if Tb < 0 or Tb > T:
            status[b] = STATUS_SEQ_LEN_OUT_OF_RANGE
            loss[b] = np.inf
            continue
        l, g, s = fn(logits[:Tb, b, :], lab, blank, dtype=dtype)
        loss[b], status[b] = l, s
        grad[:Tb, b, :] = g
    return loss, grad, status


def greedy_decode(logits, seq_len, blank=None, merge_repeated=True):
    """``tf.nn.ctc_greedy_decoder`` (``networks/tfnetwork.py:63``).

    First-index argmax over RAW logits per frame, drop blank, merge repeats
    (``prev`` is updated on every frame, so ``a,blank,a`` -> ``a,a``).
    Returns ``(hyp_values i64[M], hyp_offsets i32[B+1], neg_sum_logits f32[B])``.
    ``neg_sum_logits`` is accumulated in float32 in frame order, as TF does.
    """
    logits = np.asarray(logits)
    T, B, C = logits.shape
    if blank is None:
        blank = C - 1
    vals = []
    offsets = np.zeros(B + 1, dtype=np.int32)
    nsl = np.zeros(B, dtype=np.float32)
    for b in range(B):
        Tb = int(seq_len[b])
        row = logits[:Tb, b, :]
        am = row.argmax(axis=1) if Tb else np.zeros(0, dtype=np.int64)  # first max
        mx = row.max(axis=1).astype(np.float32) if Tb else np.zeros(0, dtype=np.float32)
        acc = np.float32(0.0)
        for m in mx:
            acc = np.float32(acc - m)
        nsl[b] = acc
        prev = -1
        n = 0
        for c in am:
            c = int(c)
            if c != blank and not (merge_repeated and c == prev):
                vals.append(c)
                n += 1
            prev = c
        offsets[b + 1] = offsets[b] + n
    return np.asarray(vals, dtype=np.int64), offsets, nsl


def csr_to_sparse(values, offsets):
    """CSR hypotheses -> TF SparseTensor triple ``(indices i64[M,2], values, dense_shape i64[2])``."""
    B = len(offsets) - 1
    lens = np.diff(offsets).astype(np.int64)
    rows = np.repeat(np.arange(B, dtype=np.int64), lens)
    cols = (np.arange(int(offsets[-1]), dtype=np.int64)
            - np.repeat(np.asarray(offsets[:-1], dtype=np.int64), lens))
    indices = np.stack([rows, cols], axis=1) if rows.size else np.zeros((0, 2), np.int64)
    shape = np.asarray([B, int(lens.max()) if B and lens.size else 0], dtype=np.int64)
    return indices, np.asarray(values), shape


def levenshtein(a, b):
    """Unit-cost Levenshtein distance (``core/lib/gtl/edit_distance.h`` semantics)."""
    a = np.asarray(a)
    b = np.asarray(b)
    n, m = a.size, b.size
    if n == 0:
        return int(m)
    if m == 0:
        return int(n)
    prev = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = i
        sub = prev[:-1] + (b != a[i - 1])
        dele = prev[1:] + 1
        best = np.minimum(sub, dele)
        # insertions: cur[j] = min(best[j-1], cur[j-1] + 1)  -> running min-plus scan
        run = cur[0]
        for j in range(1, m + 1):
            run = min(best[j - 1], run + 1)
            cur[j] = run
        prev = cur
    return int(prev[m])


def edit_distance(hyp_values, hyp_offsets, truth_values, truth_offsets, normalize=True):
    """``tf.edit_distance(hyp, truth, normalize=True)`` per batch row
    (``networks/tfnetwork.py:68``).

    Returns ``(dist i32[B], ler f32[B])``.  Empty-side conventions (definition):
    hyp empty & truth non-empty -> dist=|truth|, ler=1; truth empty & hyp
    non-empty -> dist=|hyp|, ler=+inf; both empty -> 0.
    """
    B = len(hyp_offsets) - 1
    dist = np.zeros(B, dtype=np.int32)
    ler = np.zeros(B, dtype=np.float32)
    for b in range(B):
        h = np.asarray(hyp_values[hyp_offsets[b]:hyp_offsets[b + 1]]).astype(np.int32)
        t = np.asarray(truth_values[truth_offsets[b]:truth_offsets[b + 1]]).astype(np.int32)
        d = levenshtein(t, h)
        dist[b] = d
        if not normalize:
            ler[b] = np.float32(d)
        elif t.size == 0:
            ler[b] = np.float32(np.inf) if d != 0 else np.float32(0.0)
        else:
            ler[b] = np.float32(d) / np.float32(t.size)
    return dist, ler

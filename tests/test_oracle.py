"""Pins the CPU oracle (numpy + C) before anything is compared with it.

The reference ships no tests or golden vectors for this path, so the
pins are: (i) fixtures produced by torch's independent CPU CTC / torchaudio
(tests/golden/make_golden.py) and (ii) closed-form known answers.
"""
import math

import numpy as np
import pytest

from oracle import c_oracle, ctc_oracle as o

from conftest import make_batch


def test_numpy_oracle_matches_golden(golden):
    g = golden
    loss, grad, status = o.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                         g["seq_len"])
    assert (status == 0).all()
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(grad, g["grad"], rtol=0, atol=1e-10)
    hv, ho, nsl = o.greedy_decode(g["logits"], g["seq_len"])
    assert np.array_equal(hv, g["hyp_values"]) and np.array_equal(ho, g["hyp_offsets"])
    np.testing.assert_allclose(nsl, g["neg_sum_logits"], rtol=2e-6)
    dist, ler = o.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    assert np.array_equal(dist, g["dist"])


def test_scalar_and_vectorised_numpy_agree():
    g = make_batch(7, T=25, B=4, C=6, Lmax=7)
    a = o.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"],
                        vectorised=False)
    b = o.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"],
                        vectorised=True)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
    np.testing.assert_allclose(a[1], b[1], atol=1e-12)
    assert np.array_equal(a[2], b[2])


# The f32 body computes like TF's kernel (float log-space): log-values of magnitude ~|loss| carry
# an ulp of ~|loss|*6e-8, so its gradient drifts from the f64 truth by a few 1e-4 once T reaches the
# hundreds -- hence the loose bound here.  CUDA parity tests compare with the f64 body.
@pytest.mark.parametrize("precision,ltol,gtol", [("f64", 1e-9, 1e-9), ("f32", 2e-5, 1e-3)])
def test_c_oracle_matches_golden(golden, precision, ltol, gtol):
    g = golden
    loss, grad, status = c_oracle.ctc_loss_grad(g["logits"], g["label_values"],
                                                g["label_offsets"], g["seq_len"],
                                                precision=precision)
    assert (status == 0).all()
    np.testing.assert_allclose(loss, g["loss"], rtol=ltol, atol=ltol)
    np.testing.assert_allclose(grad, g["grad"], rtol=0, atol=max(gtol, 6e-8))
    hv, ho, nsl = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    assert np.array_equal(hv, g["hyp_values"]) and np.array_equal(ho, g["hyp_offsets"])
    dist, ler = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    assert np.array_equal(dist, g["dist"])
    L = np.diff(g["label_offsets"])
    for b in range(len(L)):
        if L[b]:
            assert ler[b] == np.float32(dist[b]) / np.float32(L[b])


def test_c_oracle_grad_loss_scaling_and_padding():
    g = make_batch(11, T=40, B=5, C=9, Lmax=10)
    gl = np.linspace(0.1, 1.0, 5).astype(np.float32)
    l1, g1, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                       g["seq_len"])
    l2, g2, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                       g["seq_len"], grad_loss=gl)
    np.testing.assert_allclose(g2, g1 * gl[None, :, None], atol=1e-7)
    for b in range(5):
        assert not g1[g["seq_len"][b]:, b, :].any()          # padded frames: exactly zero
        np.testing.assert_allclose(g1[:g["seq_len"][b], b, :].sum(-1), 0, atol=2e-6)


# ---------------------------------------------------------------- closed-form known answers
def test_known_T1_L0_and_L1():
    x = np.array([[[0.3, -1.2, 2.0, 0.5]]], dtype=np.float64)       # T=1,B=1,C=4, blank=3
    ls = x[0, 0] - (x[0, 0].max() + math.log(np.exp(x[0, 0] - x[0, 0].max()).sum()))
    loss, grad, st = o.ctc_loss_grad(x, np.zeros(0, np.int32), np.array([0, 0], np.int32), [1])
    assert st[0] == 0 and abs(loss[0] + ls[3]) < 1e-12
    y = np.exp(ls)
    np.testing.assert_allclose(grad[0, 0], y - np.eye(4)[3], atol=1e-12)
    loss, grad, st = o.ctc_loss_grad(x, np.array([1], np.int32), np.array([0, 1], np.int32), [1])
    assert abs(loss[0] + ls[1]) < 1e-12
    np.testing.assert_allclose(grad[0, 0], y - np.eye(4)[1], atol=1e-12)


def _count_paths(T, lab, blank):
    """Number of length-T alignments that collapse to lab (brute-force DP over l')."""
    U = 2 * len(lab) + 1
    lp = [blank if u % 2 == 0 else lab[u // 2] for u in range(U)]
    a = [0] * U
    a[0] = 1
    if U > 1:
        a[1] = 1
    for _ in range(1, T):
        n = [0] * U
        for u in range(U):
            n[u] = a[u] + (a[u - 1] if u > 0 else 0)
            if u > 1 and lp[u] != blank and lp[u] != lp[u - 2]:
                n[u] += a[u - 2]
        a = n
    return a[U - 1] + (a[U - 2] if U > 1 else 0)


@pytest.mark.parametrize("T,lab,C", [(5, [0, 1], 4), (6, [2, 2, 0], 5), (4, [], 3), (7, [1, 0, 1], 3)])
def test_known_uniform_logits_path_count(T, lab, C):
    x = np.zeros((T, 1, C))
    loss, grad, st = o.ctc_loss_grad(x, np.asarray(lab, np.int32),
                                     np.array([0, len(lab)], np.int32), [T])
    want = -math.log(_count_paths(T, lab, C - 1) / float(C) ** T)
    assert st[0] == 0 and abs(loss[0] - want) < 1e-10
    lc, _, _ = c_oracle.ctc_loss_grad(x, np.asarray(lab, np.int32),
                                      np.array([0, len(lab)], np.int32), [T])
    assert abs(lc[0] - want) < 1e-10


def test_known_infeasible_repeat_needs_three_frames():
    # l = [a, a] needs a, blank, a  -> T=2 is infeasible: loss = +inf, status flags, grad = softmax
    x = np.random.default_rng(0).normal(size=(2, 1, 3))
    loss, grad, st = o.ctc_loss_grad(x, np.array([0, 0], np.int32), np.array([0, 2], np.int32), [2])
    assert np.isinf(loss[0]) and st[0] & o.STATUS_NOT_ENOUGH_TIME and st[0] & o.STATUS_NO_VALID_PATH
    y = np.exp(x[:, 0] - x[:, 0].max(-1, keepdims=True))
    y /= y.sum(-1, keepdims=True)
    np.testing.assert_allclose(grad[:, 0], y, atol=1e-12)
    lc, gc, sc = c_oracle.ctc_loss_grad(x, np.array([0, 0], np.int32), np.array([0, 2], np.int32), [2])
    assert np.isinf(lc[0]) and sc[0] == st[0]
    np.testing.assert_allclose(gc[:, 0], y, atol=1e-6)


def test_status_flags():
    x = np.zeros((4, 3, 5), np.float32)
    vals = np.array([0, 4, 1, 2], np.int32)          # row 0 holds label 4 == blank -> invalid
    offs = np.array([0, 2, 3, 4], np.int32)
    seq = np.array([4, 9, 4], np.int32)              # row 1: seq_len > T
    for impl in (o, c_oracle):
        loss, grad, st = impl.ctc_loss_grad(x, vals, offs, seq)
        assert st[0] == o.STATUS_LABEL_OUT_OF_RANGE and st[1] == o.STATUS_SEQ_LEN_OUT_OF_RANGE
        assert st[2] == 0 and np.isfinite(loss[2])


def test_known_greedy_collapse_and_ties():
    a, b, blank = 0, 1, 3
    path = [a, a, blank, a, b, b]
    x = np.full((6, 1, 4), -1.0, np.float32)
    for t, c in enumerate(path):
        x[t, 0, c] = 2.0
    for impl in (o, c_oracle):
        hv, ho, nsl = impl.greedy_decode(x, [6])
        assert hv.tolist() == [a, a, b] and ho.tolist() == [0, 3] and nsl[0] == -12.0
        hv, ho, _ = impl.greedy_decode(x, [6], merge_repeated=False)
        assert hv.tolist() == [a, a, a, b, b]
        hv, ho, _ = impl.greedy_decode(x, [2])             # seq_len cuts the utterance
        assert hv.tolist() == [a]
    # ties: the FIRST maximal index wins (Eigen maxCoeff), including blank-vs-label ties
    x = np.zeros((2, 1, 4), np.float32)
    x[1, 0, 2] = x[1, 0, 3] = 5.0
    for impl in (o, c_oracle):
        hv, _, _ = impl.greedy_decode(x, [2])
        assert hv.tolist() == [0, 2]


def test_known_edit_distance_conventions():
    hv = np.array([1, 2, 3, 7, 7], np.int64)
    ho = np.array([0, 3, 3, 5, 5], np.int32)          # rows: [1,2,3], [], [7,7], []
    tv = np.array([1, 3, 4, 5], np.int32)
    to = np.array([0, 2, 4, 4, 4], np.int32)          # rows: [1,3], [4,5], [], []
    for impl in (o, c_oracle):
        d, ler = impl.edit_distance(hv, ho, tv, to)
        assert d.tolist() == [1, 2, 2, 0]
        assert ler[0] == 0.5 and ler[1] == 1.0 and np.isinf(ler[2]) and ler[3] == 0.0
        d, ler = impl.edit_distance(hv, ho, tv, to, normalize=False)
        assert ler.tolist() == [1.0, 2.0, 2.0, 0.0]
    assert o.levenshtein(list(b"kitten"), list(b"sitting")) == 3


def test_sparse_csr_roundtrip():
    idx = np.array([[0, 0], [0, 1], [2, 0]], np.int64)
    vals = np.array([5, 6, 7], np.int32)
    v, off = o.sparse_to_csr((idx, vals, np.array([3, 2], np.int64)))
    assert off.tolist() == [0, 2, 2, 3] and v.tolist() == [5, 6, 7]
    i2, v2, s2 = o.csr_to_sparse(v, off)
    assert np.array_equal(i2, idx) and s2.tolist() == [3, 2]
    with pytest.raises(ValueError):
        o.sparse_to_csr((idx[::-1], vals, np.array([3, 2], np.int64)))


def _tf_basic_case():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "tf_ctc_loss_op_test_basic.json")) as f:
        return json.load(f)


def test_tf_published_ctc_loss_basic_case():
    """TensorFlow's own ctc_loss_op_test.py basic case (tests/golden/tf_ctc_loss_op_test_basic.json): the oracle, numpy
    and C, gives the published losses to their six printed digits, the published gradient entries, and otherwise the
    softmax itself (TF's tables equal the input matrix wherever the class is not on the target's path at that frame)."""
    from oracle import c_oracle
    d = _tf_basic_case()
    blank = d["blank"]
    for case in d["cases"]:
        p = np.asarray(case["probs"], np.float64)
        lab = np.asarray(case["labels"], np.int32)
        loss, grad, st = o.ctc_loss_grad_one(np.log(p), lab, blank)
        assert st == 0
        assert abs(loss - case["loss"]) < 5e-6
        for t, c, g in case["grad_entries"]:
            assert abs(grad[t, c] - g) < 2e-6
        x = np.log(p).astype(np.float32)[:, None, :]
        closs, cgrad, cst = c_oracle.ctc_loss_grad(x, lab, np.asarray([0, lab.size], np.int32),
                                                   np.asarray([5], np.int32), precision="f64")
        assert cst[0] == 0 and abs(closs[0] - case["loss"]) < 5e-6
        for t, c, g in case["grad_entries"]:
            assert abs(cgrad[t, 0, c] - g) < 2e-6


def tf_greedy_case():
    """-> (fixture, logits float32 [6, 2, 4]) of TensorFlow's published greedy decoder test."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "tf_ctc_decoder_ops_test_greedy.json")) as f:
        g = json.load(f)
    with np.errstate(divide="ignore"):
        x = np.log(np.stack([np.asarray(g["input_prob_matrix_0"], np.float32),
                             np.asarray(g["input_prob_matrix_1"], np.float32)], axis=1))
    return g, np.ascontiguousarray(x.astype(np.float32))


def test_tf_published_greedy_decoder_case():
    """TensorFlow's own ctc_decoder_ops_test.py greedy case (tests/golden/tf_ctc_decoder_ops_test_greedy.json): both
    oracles return the published SparseTensor and log probabilities (-inf logits included)."""
    g, x = tf_greedy_case()
    want_lp = np.array([np.sum(-np.log(np.asarray(p, np.float32))) for p in g["max_probs"]], np.float32)
    for dec in (o.greedy_decode, c_oracle.greedy_decode):
        vals, offs, nsl = dec(x, np.asarray(g["seq_len"], np.int32))
        assert vals.tolist() == g["values"]
        idx, _, shape = o.csr_to_sparse(vals, offs)
        assert np.asarray(idx).tolist() == g["indices"] and np.asarray(shape).tolist() == g["dense_shape"]
        assert offs.tolist() == [0, 2, 5]
        assert np.allclose(nsl, want_lp, rtol=1e-6)


def test_tf_documented_edit_distance_example():
    """The example in tf.edit_distance's own docstring (normalize=True): hypothesis (0,0)=[a], (1,0)=[b]; truth
    (0,1)=[a], (1,0)=[b,c], (1,1)=[a] -> [[inf, 1.0], [0.5, 1.0]].  The rank-3 tensor flattened to four utterances
    (a, b, c = 0, 1, 2); numpy and C oracle."""
    hyp_vals, hyp_offs = np.array([0, 1], np.int64), np.array([0, 1, 1, 2, 2], np.int32)
    tr_vals, tr_offs = np.array([0, 1, 2, 0], np.int32), np.array([0, 0, 1, 3, 4], np.int32)
    for ed in (o.edit_distance, c_oracle.edit_distance):
        d, ler = ed(hyp_vals, hyp_offs, tr_vals, tr_offs)
        assert np.asarray(d).tolist() == [1, 1, 1, 1]
        ler = np.asarray(ler, np.float64)
        assert np.isinf(ler[0]) and ler[1] == 1.0 and ler[2] == 0.5 and ler[3] == 1.0

"""ctypes binding of oracle/ctc_oracle.c (test infrastructure only).

Builds ``liboracle_ctc.so`` with ``make`` on first use if it is missing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle_ctc.so")
    srcs = [os.path.join(_HERE, f) for f in ("ctc_oracle.c", "beam_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_ctc.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def num_threads():
    return int(lib().oracle_num_threads())


def ctc_loss_grad(logits, label_values, label_offsets, seq_len, blank=None, grad_loss=None,
                  precision="f64", want_grad=True, num_threads=0):
    """Returns (loss[B], grad[T,B,C] f32 or None, status[B])."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else int(blank)
    lv = np.ascontiguousarray(label_values, dtype=np.int32)
    lo = np.ascontiguousarray(label_offsets, dtype=np.int32)
    sl = np.ascontiguousarray(seq_len, dtype=np.int32)
    real = np.float64 if precision == "f64" else np.float32
    creal = ctypes.c_double if precision == "f64" else ctypes.c_float
    loss = np.zeros(B, dtype=real)
    grad = np.empty((T, B, C), dtype=np.float32) if want_grad else None
    gl = None if grad_loss is None else np.ascontiguousarray(grad_loss, dtype=np.float32)
    status = np.zeros(B, dtype=np.int32)
    fn = getattr(lib(), "oracle_ctc_loss_grad_" + precision)
    fn.restype = ctypes.c_int
    fn(_p(logits, ctypes.c_float), T, B, C, _p(lv, ctypes.c_int32), _p(lo, ctypes.c_int32),
       _p(sl, ctypes.c_int32), blank, _p(loss, creal), _p(grad, ctypes.c_float),
       _p(gl, ctypes.c_float), _p(status, ctypes.c_int32), int(num_threads))
    return loss, grad, status


def greedy_decode(logits, seq_len, blank=None, merge_repeated=True, num_threads=0):
    """Returns (hyp_values i64[M], hyp_offsets i32[B+1], neg_sum_logits f32[B])."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else int(blank)
    sl = np.ascontiguousarray(seq_len, dtype=np.int32)
    hyp = np.zeros((B, max(T, 1)), dtype=np.int64)
    hl = np.zeros(B, dtype=np.int32)
    nsl = np.zeros(B, dtype=np.float32)
    lib().oracle_greedy_decode(_p(logits, ctypes.c_float), T, B, C, _p(sl, ctypes.c_int32), blank,
                               int(bool(merge_repeated)), _p(hyp, ctypes.c_int64),
                               _p(hl, ctypes.c_int32), _p(nsl, ctypes.c_float), int(num_threads))
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(hl)
    vals = (np.concatenate([hyp[b, :hl[b]] for b in range(B)]) if B else np.zeros(0, np.int64))
    return vals.astype(np.int64), offs, nsl


def edit_distance(hyp_values, hyp_offsets, truth_values, truth_offsets, normalize=True,
                  num_threads=0):
    hv = np.ascontiguousarray(hyp_values, dtype=np.int32)
    ho = np.ascontiguousarray(hyp_offsets, dtype=np.int32)
    tv = np.ascontiguousarray(truth_values, dtype=np.int32)
    to = np.ascontiguousarray(truth_offsets, dtype=np.int32)
    B = ho.size - 1
    dist = np.zeros(B, dtype=np.int32)
    ler = np.zeros(B, dtype=np.float32)
    lib().oracle_edit_distance(_p(hv, ctypes.c_int32), _p(ho, ctypes.c_int32),
                               _p(tv, ctypes.c_int32), _p(to, ctypes.c_int32), B,
                               int(bool(normalize)), _p(dist, ctypes.c_int32),
                               _p(ler, ctypes.c_float), int(num_threads))
    return dist, ler


def beam_search(logits, seq_len, beam_width=100, top_paths=1, merge_repeated=True, blank=None, num_threads=0,
                with_margin=False):
    """oracle/beam_oracle.c (TF's trie + TopN control flow).  Returns (hyp i64[B,P,T], hyp_len i32[B,P],
    log_prob f64[B,P]); logits may be any [T,B,C] float32 view with a dense class axis.  ``with_margin`` adds
    margin f64[B,2]: [:, 0] the smallest NONZERO difference of totals over all decisions the search took for that
    utterance — a result whose margin is within rounding of zero hangs on the last bit of exp/log — and [:, 1] the
    number of decisions between bitwise-equal totals (twins from equal logits in a frame: deterministic everywhere;
    with quantised logits also accidental equalities that another libm may break)."""
    logits = np.asarray(logits, dtype=np.float32)
    if logits.strides[2] != 4:
        logits = np.ascontiguousarray(logits)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else int(blank)
    sl = np.ascontiguousarray(seq_len, dtype=np.int32)
    P = int(top_paths)
    hyp = np.zeros((B, P, max(T, 1)), dtype=np.int64)
    hl = np.zeros((B, P), dtype=np.int32)
    lp = np.zeros((B, P), dtype=np.float64)
    margin = np.full((B, 2), np.inf, dtype=np.float64)
    fn = lib().oracle_beam_search_margin
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    fn(logits.ctypes.data, T, B, C, logits.strides[0] // 4, logits.strides[1] // 4, sl.ctypes.data, blank,
       int(beam_width), P, int(bool(merge_repeated)), hyp.ctypes.data, hl.ctypes.data, lp.ctypes.data,
       margin.ctypes.data, int(num_threads))
    out = (hyp[:, :, :T] if T else hyp[:, :, :0], hl, lp)
    return out + (margin,) if with_margin else out

"""Label sparse format helpers (host side).

``sparse_tuple_from`` mirrors the reference's ``utils.py:44-58`` in name, arguments and result:
dense padded labels ``[B, Lmax]`` + ``output_lengths`` -> ``(indices i64[N,2], values i32[N],
shape i64[2])`` with ``shape = [B, max length]``.  Built with array operations rather than Python
lists; an all-empty batch raises ``ValueError`` (the reference's ``.max(0)`` on an empty array
raises there too).
"""
from __future__ import annotations

import numpy as np


def sparse_tuple_from(sequences, output_lengths):
    lens = np.asarray(output_lengths, dtype=np.int64).reshape(-1)
    B = len(sequences)
    if lens.size != B:
        raise ValueError("sparse_tuple_from: %d sequences but %d lengths" % (B, lens.size))
    if B == 0 or lens.max(initial=0) <= 0:
        raise ValueError("sparse_tuple_from: zero-size array to reduction operation maximum "
                         "(all label sequences are empty)")
    rows = np.repeat(np.arange(B, dtype=np.int64), lens)
    starts = np.cumsum(lens) - lens
    cols = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(starts, lens)
    indices = np.stack([rows, cols], axis=1)
    values = np.concatenate([np.asarray(seq)[:l] for seq, l in zip(sequences, lens)]).astype(np.int32)
    shape = np.asarray([B, int(cols.max()) + 1], dtype=np.int64)
    return indices, values, shape


def sparse_to_csr(labels):
    """``(indices, values, shape)`` -> ``(values i32[N], offsets i32[B+1], max_len)``.

    Accepts exactly what ``sparse_tuple_from`` / a TF ``SparseTensorValue`` holds; rows must be
    batch-ordered (TF: "indices ordered by batch"), which makes offsets a prefix sum of row counts.
    """
    indices, values, shape = labels
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    values = np.ascontiguousarray(np.asarray(values).reshape(-1), dtype=np.int32)
    B = int(np.asarray(shape).reshape(-1)[0])
    if indices.shape[0] != values.shape[0]:
        raise ValueError("labels: indices has %d rows but values has %d" % (indices.shape[0], values.shape[0]))
    rows = indices[:, 0]
    if rows.size:
        if rows.min() < 0 or rows.max() >= B:
            raise ValueError("labels: batch index outside [0, %d)" % B)
        if np.any(np.diff(rows) < 0):
            raise ValueError("labels: indices are not ordered by batch")
    counts = np.bincount(rows, minlength=B)
    offsets = np.zeros(B + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    return values, offsets, int(counts.max()) if B else 0


def split_labels(labels, num_split):
    """``tf.sparse_split(sp_input=labels, num_split=n, axis=0)`` for the label triple
    (reference ``networks/tfnetwork.py:97-99``): equal contiguous row blocks, row indices re-based,
    every part keeps the parent's second dimension."""
    indices, values, shape = labels
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    values = np.asarray(values)
    B, width = int(shape[0]), int(shape[1])
    if B % num_split:
        raise ValueError("split_labels: batch %d not divisible by %d" % (B, num_split))
    per = B // num_split
    out = []
    for r in range(num_split):
        m = (indices[:, 0] >= r * per) & (indices[:, 0] < (r + 1) * per)
        idx = indices[m].copy()
        idx[:, 0] -= r * per
        out.append((idx, values[m], np.asarray([per, width], dtype=np.int64)))
    return out

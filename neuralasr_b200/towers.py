"""Utterance sharding over GPUs and the one collective of the path.

The reference's only parallelism is in-graph data-parallel "towers" (``networks/tfnetwork.py:88-113``):
``tf.split`` / ``tf.sparse_split(axis=0)`` cut the batch into equal contiguous blocks of utterances, each
GPU runs the whole model tail on its block, and the step's scalars are the mean of the tower means
(``tfnetwork.py:113,135-136``).  Here a tower is a process (one per GPU, ``torch.distributed``):
every rank owns a local contiguous ``[T, B/G, C]`` logits tensor, its rows of the label triple re-based to
local batch indices, and its ``seq_len``.  Losses, gradients, hypotheses and per-utterance error rates
never leave their GPU; the only exchange is one all-reduce(sum) of the float64 4-vector
``[sum loss, sum ler, sum edit distance, utterance count]`` that ``common.batch_sums`` builds on the
device — 32 bytes, latency-bound, so it is a plain NCCL all-reduce rather than a fused kernel; issued
asynchronously it overlaps the next step's kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .utils import split_labels


def tune_nccl_for_scalars():
    """Call before the NCCL communicator is created: the path's one collective is 32 bytes, one channel (one CTA) is
    all it needs, and every further NCCL CTA is a CTA slot the next step's loss kernel does not get (measured on
    8 B200: 0.317 -> 0.314 ms per step at cfg3, 97 % of linear).  Respects values already in the environment;
    ``NASR_NCCL_TUNE=0`` leaves NCCL's defaults."""
    import os
    if os.environ.get("NASR_NCCL_TUNE", "1") != "0":
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")


def shard_range(batch, rank, world):
    """Rows ``[lo, hi)`` of a global batch that tower ``rank`` owns (``tf.split`` semantics: equal blocks;
    the reference guarantees divisibility, ``config.py:35-36``)."""
    if batch % world:
        raise ValueError("batch %d is not divisible by %d towers" % (batch, world))
    per = batch // world
    return rank * per, (rank + 1) * per


def shard_inputs(labels, seq_len, rank, world, logits=None):
    """This tower's ``(labels triple, seq_len[, logits])`` cut from host-side global inputs.
    ``logits`` (numpy ``[T, B, C]``), if given, is sliced on the batch axis and made contiguous — in a real
    step each tower's model produces its local logits directly and nothing is sliced."""
    seq_len = np.asarray(seq_len)
    lo, hi = shard_range(seq_len.shape[0], rank, world)
    part = split_labels(labels, world)[rank]
    out = (part, seq_len[lo:hi].copy())
    if logits is not None:
        out = out + (np.ascontiguousarray(logits[:, lo:hi, :]),)
    return out


def all_reduce_sums(sums, group=None, async_op=False):
    """Sum the ``[sum loss, sum ler, sum dist, count]`` vectors of all towers in place (no-op without an
    initialised process group).  Works on CUDA tensors over NCCL and on CPU tensors over gloo.

    ``async_op=True`` returns ``(sums, work)``: the reduction is ordered after what the current stream has
    enqueued but the stream does not wait for it, so the next step's kernels start while the 32 bytes travel
    (the scalars are only read for logging, ``train.py``); call ``work.wait()`` — ``work`` is ``None`` without a
    process group — before reading ``sums``."""
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (sums, work) if async_op else sums


def step_scalars(sums):
    """``(mean loss, mean label error rate, total edit distance, utterances)`` from a reduced 4-vector.
    With equal shards the means equal the reference's mean of tower means (``tfnetwork.py:135-136``)."""
    s = sums.detach().to("cpu", torch.float64).numpy()
    n = max(s[3], 1.0)
    return float(s[0] / n), float(s[1] / n), int(round(s[2])), int(round(s[3]))

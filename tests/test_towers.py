"""Utterance sharding + the scalar all-reduce, world_size 2 over gloo on CPU (the N>1 host path).

The per-utterance numbers a rank contributes are produced by the oracle here (there is no GPU in this
test); what is under test is the sharding (tf.split / tf.sparse_split semantics, networks/tfnetwork.py:93-101)
and the reduction (mean of tower means, tfnetwork.py:135-136)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import make_batch  # noqa: E402
from neuralasr_b200 import towers, utils  # noqa: E402
from oracle import c_oracle  # noqa: E402


def _triple(g):
    offs = g["label_offsets"]
    B = offs.size - 1
    lens = np.diff(offs)
    rows = np.repeat(np.arange(B), lens)
    cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
    return (np.stack([rows, cols], 1).astype(np.int64), g["label_values"].astype(np.int32),
            np.asarray([B, max(int(lens.max()), 1)], np.int64))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, g, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        labels, seq, x = towers.shard_inputs(_triple(g), g["seq_len"], rank, world, logits=g["logits"])
        vals, offs, _ = utils.sparse_to_csr(labels)
        loss, _, _ = c_oracle.ctc_loss_grad(x, vals, offs, seq, precision="f64", want_grad=False)
        hv, ho, _ = c_oracle.greedy_decode(x, seq)
        d, ler = c_oracle.edit_distance(hv, ho, vals, offs)
        sums = torch.tensor([loss.sum(), ler.astype(np.float64).sum(), float(d.sum()), float(len(seq))],
                            dtype=torch.float64)
        towers.all_reduce_sums(sums)
        out[rank] = towers.step_scalars(sums) + (loss.tolist(),)
    finally:
        dist.destroy_process_group()


def test_two_towers_reduce_to_the_single_tower_result():
    g = make_batch(11, T=60, B=8, C=12, Lmax=10, mode="ragged", empty_row=False)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), g, out), nprocs=world, join=True)
    # single tower
    loss, _, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"],
                                        precision="f64", want_grad=False)
    hv, ho, _ = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    d, ler = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    for rank in range(world):
        mean_loss, mean_ler, tot_d, n = out[rank][:4]
        assert n == 8 and tot_d == int(d.sum())
        assert abs(mean_loss - loss.mean()) < 1e-9 * abs(loss.mean())
        assert abs(mean_ler - ler.astype(np.float64).mean()) < 1e-9
    # per-utterance results are untouched by sharding: concatenated shards == the global batch
    assert np.allclose(out[0][4] + out[1][4], loss, rtol=1e-12)
    # and equal shards make the global mean the mean of the tower means (tfnetwork.py:135-136)
    tower_means = [np.mean(out[r][4]) for r in range(world)]
    assert abs(np.mean(tower_means) - loss.mean()) < 1e-9 * abs(loss.mean())


def test_shard_inputs_semantics():
    g = make_batch(12, T=30, B=6, C=9, Lmax=5, mode="full", empty_row=True)
    trip = _triple(g)
    parts = [towers.shard_inputs(trip, g["seq_len"], r, 3, logits=g["logits"]) for r in range(3)]
    assert [p[2].shape for p in parts] == [(30, 2, 9)] * 3 and all(p[2].flags["C_CONTIGUOUS"] for p in parts)
    vals = np.concatenate([utils.sparse_to_csr(p[0])[0] for p in parts])
    assert np.array_equal(vals, g["label_values"])
    assert all(p[0][0][:, 0].max(initial=0) < 2 for p in parts)        # row indices re-based per tower
    with pytest.raises(ValueError):
        towers.shard_range(7, 0, 2)
    assert towers.all_reduce_sums(torch.ones(4, dtype=torch.float64)).tolist() == [1, 1, 1, 1]   # no group: no-op

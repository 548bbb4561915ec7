"""CPU tests of the host-side logic and of the C-ABI surface (no kernels are launched)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

from neuralasr_b200 import _build, _lib, utils


def _reference_sparse_tuple_from(sequences, output_lengths):
    """Behavioural restatement of reference utils.py:44-58 used only to check ours against."""
    indices, values = [], []
    for n, seq in enumerate(sequences):
        l = output_lengths[n]
        indices.extend(zip([n] * l, range(l)))
        values.extend(seq[:l])
    indices = np.asarray(indices, dtype=np.int64)
    values = np.asarray(values, dtype=np.int32)
    shape = np.asarray([len(sequences), indices.max(0)[1] + 1], dtype=np.int64)
    return indices, values, shape


def test_sparse_tuple_from_matches_reference_behaviour():
    rng = np.random.default_rng(0)
    dense = rng.integers(1, 30, size=(6, 9))
    lens = np.array([9, 0, 3, 1, 7, 0])
    a = utils.sparse_tuple_from(dense, lens)
    b = _reference_sparse_tuple_from(dense, lens)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    with pytest.raises(ValueError):
        utils.sparse_tuple_from(dense, np.zeros(6, int))   # the reference raises on an all-empty batch too


def test_sparse_to_csr_and_split():
    dense = np.arange(40).reshape(4, 10)
    lens = np.array([2, 0, 10, 5])
    trip = utils.sparse_tuple_from(dense, lens)
    vals, offs, mx = utils.sparse_to_csr(trip)
    assert offs.tolist() == [0, 2, 2, 12, 17] and mx == 10 and vals.dtype == np.int32
    assert vals[2:12].tolist() == dense[2].tolist()
    parts = utils.split_labels(trip, 2)          # tf.sparse_split(axis=0) semantics
    v0, o0, _ = utils.sparse_to_csr(parts[0])
    v1, o1, _ = utils.sparse_to_csr(parts[1])
    assert o0.tolist() == [0, 2, 2] and o1.tolist() == [0, 10, 15]
    assert parts[1][0][:, 0].min() == 0 and parts[1][2].tolist() == [2, 10]
    assert np.array_equal(np.concatenate([v0, v1]), vals)
    with pytest.raises(ValueError):
        utils.split_labels(trip, 3)
    bad = (trip[0][::-1].copy(), trip[1], trip[2])
    with pytest.raises(ValueError):
        utils.sparse_to_csr(bad)


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build_library()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "nasr_ctc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(nasr_[a-z0-9_]+)\s*\(", header))
    assert {"nasr_ctc_loss_grad_f32", "nasr_ctc_greedy_decode_i64", "nasr_edit_distance_i64"} <= declared
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "binding and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()
    assert lib.nasr_abi_version() == _lib.ABI_VERSION


def test_argument_errors_without_gpu():
    lib = _lib.load()
    n = ctypes.c_size_t()
    assert lib.nasr_ctc_workspace_bytes(1000, 256, 38, 200, ctypes.byref(n)) == _lib.OK
    # checkpoint rows: B * ceil(T/16) * 402 doubles  (DESIGN.md "workspace")
    assert n.value >= 256 * 63 * 402 * 8
    assert lib.nasr_ctc_workspace_bytes(10, 1, 0, 2, ctypes.byref(n)) == _lib.ERR_INVALID_ARGUMENT
    assert b"bad shape" in lib.nasr_last_error()
    assert lib.nasr_ctc_workspace_bytes(100, 4, 38, 100000, ctypes.byref(n)) == _lib.ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        _lib.check(lib.nasr_ctc_loss_grad_f32(None, 4, 2, 5, None, None, 1, None, 9, None, None, None,
                                              None, None, 0, None), "loss")   # blank outside [0,C)


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    from neuralasr_b200.networks import common
    x = torch.zeros(4, 2, 5)
    lab = utils.sparse_tuple_from([[1, 2], [3, 0]], [2, 1])
    with pytest.raises(ValueError, match="no CPU path"):
        common.loss(x, lab, [4, 4])
    with pytest.raises(ValueError, match="no CPU path"):
        common.decoding(x, [4, 4])
    src = open(os.path.join(ROOT, "neuralasr_b200", "networks", "common.py")).read()
    assert "oracle" not in src


def test_symbols_table_and_convert_to_str(tmp_path):
    from neuralasr_b200.symbols import Symbols
    s = Symbols(label_context=0)
    s.insert_padding()                              # preprocess_mfcc.py:81: padding first ...
    for ch in "abc_":
        s.insert_sym(ch)
    s.insert_blank()                                # ... blank last (preprocess_mfcc.py:92)
    assert s.get_padding_id() == 0 and s.get_blank_id() == s.num_classes - 1 == 5
    assert s.insert_sym("b") == 2                   # idempotent
    assert s.convert_to_str([1, 2, 4, 3, 5, 1]) == "ab ca"
    path = str(tmp_path / "symbols.txt")
    s.write(path)
    r = Symbols(0, path)
    assert r.sym_to_id == s.sym_to_id and r.counter == s.counter and r.get_sym(4) == "_"
    # label context 1: symbols are trigrams, the middle character is the label (preprocess_mfcc.py:18-29)
    t = Symbols(label_context=1)
    ids = [t.insert_sym(x) for x in ("^he", "hel", "el_", "l_o")]
    assert t.convert_to_str(ids) == "hel "


def test_dense_to_sparse_drops_eos():
    pytest.importorskip("torch")
    from neuralasr_b200.steps import dense_to_sparse
    idx, vals, shape = dense_to_sparse(np.array([[3, 0, 4], [0, 0, 0], [7, 8, 0]]))
    assert idx.tolist() == [[0, 0], [0, 2], [2, 0], [2, 1]] and vals.tolist() == [3, 4, 7, 8]
    assert shape.tolist() == [3, 3]


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: one JSON line with the
    contract's keys, the C port's rate as value."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "cfg1"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ctc_loss_grad_frames_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    for key in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config"):
        assert key in d

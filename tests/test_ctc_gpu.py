"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the golden fixtures.

Tolerances are the north star's: loss 1e-4 relative, gradient 1e-4 absolute (fp32 outputs compared
with the float64 oracle), decoded sequences / hypothesis lengths / integer edit distances bit-exact.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import c_oracle, ctc_oracle as o  # noqa: E402  (the checker, never the thing under test)

from conftest import make_batch  # noqa: E402

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_ATOL = 1e-4


@pytest.fixture(scope="module")
def common():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from neuralasr_b200.networks import common as c
    return c


@pytest.fixture(autouse=True, params=["f64", "f32"])
def narrow_kernel(request, monkeypatch):
    """Every test of this module runs with both kernels for narrow vocabularies: the fp64 kernel (csrc/ctc_fast.cu,
    the default) and the float32 one (csrc/ctc_narrow.cu, NASR_NARROW_F32=1: the library reads the variable at
    every call)."""
    if request.param == "f32":
        monkeypatch.setenv("NASR_NARROW_F32", "1")
    else:
        monkeypatch.delenv("NASR_NARROW_F32", raising=False)
    return request.param


def _triple(g):
    from neuralasr_b200.utils import sparse_to_csr  # noqa: F401
    offs = g["label_offsets"]
    B = offs.size - 1
    lens = np.diff(offs)
    rows = np.repeat(np.arange(B), lens)
    cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
    return (np.stack([rows, cols], 1).astype(np.int64), g["label_values"].astype(np.int32),
            np.asarray([B, max(int(lens.max()), 1)], np.int64))


def _run_loss(common, g, grad_loss=None, use_dlpack=False):
    x = torch.from_numpy(g["logits"]).cuda()
    gl = None if grad_loss is None else torch.from_numpy(grad_loss).cuda()
    loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"], grad_loss=gl,
                                                  use_dlpack=use_dlpack)
    torch.cuda.synchronize()
    return loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()


def _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status=None):
    if want_status is not None:
        assert np.array_equal(status, want_status)
    fin = np.isfinite(want_loss)
    assert np.array_equal(np.isfinite(loss), fin)
    np.testing.assert_allclose(loss[fin], want_loss[fin], rtol=LOSS_RTOL, atol=1e-5)
    assert np.isfinite(grad).all()
    assert np.abs(grad - want_grad).max() <= GRAD_ATOL


def test_loss_grad_matches_golden(common, golden):
    loss, grad, status = _run_loss(common, golden)
    _assert_loss_grad(loss, grad, status, golden["loss"], golden["grad"], np.zeros_like(status))
    for b, tb in enumerate(golden["seq_len"]):
        assert not grad[tb:, b, :].any()            # padded frames: exactly zero


@pytest.mark.parametrize("name,kw", [
    ("cfg1", dict(T=500, B=16, C=38, Lmax=100, mode="ragged")),
    ("cfg2", dict(T=800, B=64, C=38, Lmax=150, mode="ragged")),
    ("tight", dict(T=120, B=9, C=38, Lmax=50, mode="tight")),
    ("peaky", dict(T=400, B=12, C=38, Lmax=60, mode="ragged", peaky=True)),
    ("tiny_vocab", dict(T=64, B=7, C=2, Lmax=20, mode="ragged", repeat_p=0.0)),
    ("wide_vocab", dict(T=200, B=6, C=1024, Lmax=40, mode="ragged")),
    ("long", dict(T=1500, B=3, C=38, Lmax=300, mode="full")),
    ("odd_sizes", dict(T=17, B=3, C=5, Lmax=3, mode="ragged")),
])
def test_loss_grad_matches_oracle(common, name, kw):
    g = make_batch(hash(name) % 1000, **kw)
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    loss, grad, status = _run_loss(common, g)
    _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status)
    valid = grad[: g["seq_len"].min()]
    assert np.abs(valid.sum(-1)).max() < 1e-5       # softmax - occupancy sums to 0 over classes


def test_grad_loss_scaling_dlpack_and_repeatability(common):
    g = make_batch(5, T=90, B=8, C=38, Lmax=20)
    gl = np.linspace(0.05, 1.0, 8).astype(np.float32)
    l0, g0, s0 = _run_loss(common, g)
    l1, g1, s1 = _run_loss(common, g, grad_loss=gl)
    np.testing.assert_allclose(g1, g0 * gl[None, :, None], atol=1e-6)
    l2, g2, s2 = _run_loss(common, g, use_dlpack=True)
    assert np.array_equal(l0, l2) and np.array_equal(g0, g2) and np.array_equal(s0, s2)
    l3, g3, _ = _run_loss(common, g)
    assert np.array_equal(l0, l3) and np.array_equal(g0, g3)      # deterministic, bit for bit


def test_status_flags_and_infeasible(common):
    C = 6
    x = np.random.default_rng(3).normal(size=(5, 5, C)).astype(np.float32)
    labs = [[0, 1], [5, 1], [2, 2, 2, 2], [1], []]     # row1: label == blank; row2 needs 7 frames
    lens = [len(l) for l in labs]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    vals = np.concatenate([np.asarray(l, np.int32) for l in labs])
    seq = np.array([5, 5, 5, 9, 0], np.int32)          # row3: seq_len > T; row4: empty utterance
    g = dict(logits=x, label_values=vals, label_offsets=offs, seq_len=seq)
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(x, vals, offs, seq)
    loss, grad, status = _run_loss(common, g)
    assert status.tolist() == want_status.tolist() == [0, 1, 4 | 8, 2, 0]
    _assert_loss_grad(loss, grad, status, want_loss, want_grad)
    assert loss[4] == 0.0 and not grad[:, 4].any()
    with pytest.raises(ValueError, match="Not enough time|non-null label|sequence_length"):
        common.loss(torch.from_numpy(x).cuda(), _triple(g), seq)


def test_drop_in_loss_autograd_mean(common):
    g = make_batch(21, T=60, B=4, C=11, Lmax=9, empty_row=False)
    x = torch.from_numpy(g["logits"]).cuda().requires_grad_(True)
    out = common.loss(x, _triple(g), torch.from_numpy(g["seq_len"]))
    (out * 3.0).backward()
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"],
                                                     g["label_offsets"], g["seq_len"])
    assert abs(float(out) - want_loss.mean()) <= LOSS_RTOL * want_loss.mean()
    np.testing.assert_allclose(out.per_utterance.cpu().numpy(), want_loss, rtol=LOSS_RTOL)
    assert np.abs(x.grad.cpu().numpy() - 3.0 * want_grad / 4).max() <= GRAD_ATOL


@pytest.mark.parametrize("kw", [
    dict(T=800, B=64, C=38, Lmax=150, mode="ragged", peaky=True),
    dict(T=300, B=5, C=38, Lmax=60, mode="ragged"),
    dict(T=2500, B=3, C=7, Lmax=40, mode="full"),
    dict(T=33, B=4, C=1024, Lmax=8, mode="ragged"),
])
def test_decode_and_ler_match_oracle(common, kw):
    g = make_batch(31, **kw)
    x = torch.from_numpy(g["logits"]).cuda()
    decoded, nsl = common.decoding(x, g["seq_len"])
    hv, ho, want_nsl = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    idx, vals, shape = decoded
    wi, wv, ws = o.csr_to_sparse(hv, ho)
    assert np.array_equal(vals.cpu().numpy(), wv) and vals.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(shape.cpu().numpy(), ws)
    assert np.array_equal(decoded.hyp_len.cpu().numpy(), np.diff(ho))
    assert nsl.shape == (kw["B"], 1)
    assert np.array_equal(nsl.cpu().numpy()[:, 0], want_nsl)       # same fp32 accumulation order
    want_d, want_ler = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    mean = common.label_error_rate(decoded, _triple(g))
    assert np.array_equal(mean.distances.cpu().numpy(), want_d)
    assert np.array_equal(mean.per_utterance.cpu().numpy(), want_ler)
    fin = np.isfinite(want_ler)
    if fin.all():
        assert abs(float(mean) - want_ler.astype(np.float64).mean()) < 1e-6
    else:
        assert np.isinf(float(mean))
    # same metric from the raw SparseTensor triple (the form LAS's caller holds, las.py:116-117)
    mean2 = common.label_error_rate((idx, vals, shape), _triple(g))
    assert np.array_equal(mean2.distances.cpu().numpy(), want_d)


def test_greedy_decode_layouts_and_ties(common):
    """Narrow rows against the oracle: even and odd row lengths, a batch-major view, rows offset by one float (no
    8-byte alignment), more frames than one staging pass, quantised logits (ties: first index wins)."""
    rng = np.random.default_rng(9)
    for C, T, B, layout in [(38, 333, 21, 0), (37, 100, 5, 0), (64, 47, 33, 1), (2, 40, 3, 0), (41, 130, 17, 2),
                            (38, 2100, 2, 1), (1, 20, 4, 0)]:
        x = np.round(rng.normal(size=(T, B, C)) * 2).astype(np.float32) / 2
        seq = rng.integers(0, T + 1, B).astype(np.int32)
        seq[0] = T
        xt = torch.from_numpy(x).cuda()
        if layout == 1:
            xt = xt.transpose(0, 1).contiguous().transpose(0, 1)
        elif layout == 2:
            big = torch.zeros((T, B, C + 3), device="cuda")
            big[:, :, 1:C + 1] = xt
            xt = big[:, :, 1:C + 1]
        hv, ho, want_nsl = c_oracle.greedy_decode(x, seq)
        dec, nsl = common.decoding(xt, seq)
        assert np.array_equal(dec.values.cpu().numpy(), hv), (C, T, B, layout)
        assert np.array_equal(dec.hyp_len.cpu().numpy(), np.diff(ho))
        assert np.array_equal(nsl.cpu().numpy()[:, 0], want_nsl)


def test_decode_known_cases(common):
    a, b, blank = 0, 1, 3
    x = np.full((6, 2, 4), -1.0, np.float32)
    for t, c in enumerate([a, a, blank, a, b, b]):
        x[t, 0, c] = 2.0
    x[:, 1, :] = 0.0
    x[1, 1, 2] = x[1, 1, 3] = 5.0                       # tie between label 2 and blank: first index wins
    xt = torch.from_numpy(x).cuda()
    dec, nsl = common.decoding(xt, [6, 2])
    idx, vals, shape = dec
    assert vals.tolist() == [a, a, b, 0, 2] and shape.tolist() == [2, 3]
    assert idx.tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1]]
    assert nsl[:, 0].tolist() == [-12.0, -5.0]
    dec2, _ = common.decoding(xt, [6, 2], merge_repeated=False)
    assert dec2.values.tolist() == [a, a, a, b, b, 0, 2]
    # edit-distance conventions: empty hypothesis -> 1.0; empty truth with non-empty hypothesis -> inf
    from neuralasr_b200.networks.common import edit_distance
    hyp = (np.array([[0, 0], [0, 1], [0, 2], [2, 0], [2, 1]], np.int64), np.array([1, 2, 3, 7, 7], np.int64),
           np.array([4, 3], np.int64))
    truth = (np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.int64), np.array([1, 3, 4, 5], np.int32),
             np.array([4, 2], np.int64))
    d, ler = edit_distance(hyp, truth)
    assert d.tolist() == [1, 2, 2, 0]
    ler = ler.cpu().numpy()
    assert ler[0] == 0.5 and ler[1] == 1.0 and np.isinf(ler[2]) and ler[3] == 0.0


def test_host_buffer_context_matches_device_path(common):
    import ctypes
    from neuralasr_b200 import host
    g = make_batch(77, T=100, B=8, C=38, Lmax=25)
    ctx = host.HostContext(0, 100, 8, 38, 25)
    out = ctx.step(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"])
    loss, grad, status = _run_loss(common, g)
    assert np.array_equal(out["loss"], loss) and np.array_equal(out["grad"], grad)
    hv, ho, _ = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    want_d, want_ler = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    assert np.array_equal(out["hyp_len"], np.diff(ho)) and np.array_equal(out["dist"], want_d)
    assert np.array_equal(out["ler"], want_ler)
    ctx.close()


def test_full_size_properties(common):
    """BASELINE headline shape (cfg3): size-independent properties, and the whole batch against the C oracle."""
    g = make_batch(1234, T=1000, B=256, C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
    loss, grad, status = _run_loss(common, g)
    assert (status == 0).all() and np.isfinite(loss).all() and np.isfinite(grad).all()
    assert np.abs(grad.sum(-1)).max() < 2e-5            # rows of softmax - occupancy sum to zero
    # occupancy of the blank+labels is a distribution: softmax - grad lies in [0, 1]
    x = torch.from_numpy(g["logits"]).cuda()
    occ = torch.softmax(x, -1).cpu().numpy() - grad
    assert occ.min() > -1e-5 and occ.max() < 1 + 1e-5
    # expected label counts: sum_t occupancy[t, b, c != blank] >= number of label positions of c ... and
    # total non-blank occupancy is at least L (each label is emitted at least once)
    L = np.diff(g["label_offsets"])
    assert (occ[:, :, :-1].sum((0, 2)) >= L - 1e-2).all()
    # and the whole batch against the C oracle (float64; a quarter of a second on the box's host threads)
    wl, wg, ws = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    _assert_loss_grad(loss, grad, status, wl, wg, ws)


def test_full_size_cfg5_whole_batch_against_oracle(common):
    """BASELINE cfg5 (B=128, T=800, C=1024, L<=150), the bandwidth-dominated shape: every loss and every gradient
    element against the C oracle, and no utterance handed to the retry kernel."""
    g = make_batch(4321, T=800, B=128, C=1024, Lmax=150, mode="ragged", Lmin=75, empty_row=False)
    loss, grad, status = _run_loss(common, g)
    assert not _flags(common, 128).any()
    wl, wg, ws = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    _assert_loss_grad(loss, grad, status, wl, wg, ws)
    assert np.abs(grad.sum(-1)).max() < 5e-5


# ------------------------------------------------------------------------------------------------
# the two kernels behind nasr_ctc_loss_grad: throughput kernel (ctc_fast.cu) + robust retry (ctc_loss.cu)
# ------------------------------------------------------------------------------------------------
@pytest.fixture
def debug_paths(common):
    yield common.debug_config
    common.debug_config(0, 0)            # always restore the default dispatch


def _flags(common, B):
    return common.retry_flags(torch.device("cuda", 0), B).cpu().numpy()


@pytest.mark.parametrize("kw", [
    dict(T=500, B=16, C=38, Lmax=100, mode="ragged"),                 # BASELINE cfg1
    dict(T=400, B=12, C=38, Lmax=60, mode="ragged", peaky=True),
    dict(T=100, B=5, C=64, Lmax=20, mode="ragged"),                   # widest vocabulary the fast kernel takes
    dict(T=40, B=3, C=5, Lmax=6, mode="ragged"),
    dict(T=3000, B=2, C=38, Lmax=600, mode="full", empty_row=False),  # BASELINE cfg4: 20 slots per lane
    dict(T=200, B=6, C=1024, Lmax=40, mode="ragged"),                 # wide-vocabulary variant (BASELINE cfg5 classes)
    dict(T=400, B=4, C=132, Lmax=150, mode="ragged"),                 # wide variant, C not a multiple of 128
    dict(T=300, B=3, C=512, Lmax=200, mode="full", empty_row=False),
    dict(T=120, B=3, C=1028, Lmax=30, mode="ragged"),                 # just past the register-held row: streamed rows
    dict(T=150, B=4, C=3000, Lmax=60, mode="ragged", peaky=True),     # the reference's trigram vocabularies
    dict(T=90, B=2, C=6000, Lmax=20, mode="ragged"),                  # retry kernel plans 4-frame segments here
    dict(T=70, B=2, C=8192, Lmax=12, mode="ragged"),                  # widest row the throughput kernel takes
    dict(T=100, B=5, C=131, Lmax=25, mode="ragged"),                  # C % 4 != 0: rows at every misalignment, streamed
    dict(T=80, B=3, C=1001, Lmax=20, mode="ragged", peaky=True),
    dict(T=60, B=3, C=3187, Lmax=15, mode="ragged"),
    dict(T=700, B=2, C=1024, Lmax=300, mode="full", empty_row=False),  # wide with 223..318 labels: ten slots per lane
    dict(T=680, B=2, C=3000, Lmax=318, mode="ragged", empty_row=False),
])
def test_each_kernel_alone_matches_oracle(common, debug_paths, kw):
    g = make_batch(4242, **kw)
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    debug_paths(1, 0)                    # robust kernel only
    loss, grad, status = _run_loss(common, g)
    _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status)
    debug_paths(2, 0)                    # throughput kernel only: it must take every regular utterance itself
    loss, grad, status = _run_loss(common, g)
    flags = _flags(common, kw["B"])
    regular = g["seq_len"] >= 16
    assert not flags[regular].any(), "throughput kernel handed over %s" % flags
    ok = flags == 0
    np.testing.assert_allclose(loss[ok], want_loss[ok], rtol=LOSS_RTOL, atol=1e-5)
    assert np.abs(grad[:, ok] - want_grad[:, ok]).max() <= GRAD_ATOL


@pytest.mark.parametrize("split", [8, 16, 64, 104])
def test_uneven_meeting_points(common, debug_paths, split):
    """The forward/backward halves may meet anywhere: same results for lopsided splits."""
    g = make_batch(99, T=120, B=6, C=38, Lmax=30, mode="ragged")
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                                     g["seq_len"], precision="f64")
    debug_paths(2, split)
    loss, grad, status = _run_loss(common, g)
    assert not _flags(common, 6).any()
    _assert_loss_grad(loss, grad, status, want_loss, want_grad)


def test_out_of_range_inputs_go_to_the_robust_kernel(common, debug_paths):
    """Logit gaps beyond what a float ratio can hold, short utterances, infeasible and invalid rows: the
    throughput kernel must flag them (never answer wrongly) and the default path must still be right."""
    g = make_batch(7, T=64, B=8, C=12, Lmax=10, mode="full", empty_row=False)
    x = g["logits"]
    x[10:20, 0, 11] -= 60.0              # blank nearly impossible for ten frames: ratio emissions ~e^60 each
    x[30, 1, 3] += 60.0                  # one class dominates a frame by e^60
    x[:, 2, :] *= 5.0                    # large dynamic range everywhere (gaps up to ~e^80)
    g["seq_len"][3] = 9                  # shorter than two chunks
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        x, g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    debug_paths(2, 0)
    loss, grad, status = _run_loss(common, g)
    flags = _flags(common, 8)
    assert flags[3]
    # whatever the throughput kernel keeps for itself must be right (the float32 kernel's emissions are in units of
    # max(blank, largest class / 32), so a vanishing blank no longer forces a hand-over; the fp64 kernel flags it)
    ok = flags == 0
    np.testing.assert_allclose(loss[ok], want_loss[ok], rtol=LOSS_RTOL, atol=1e-5)
    assert np.abs(grad[:, ok] - want_grad[:, ok]).max() <= GRAD_ATOL
    debug_paths(0, 0)
    loss, grad, status = _run_loss(common, g)
    _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status)


def test_loss_only_call(common):
    g = make_batch(3, T=200, B=5, C=38, Lmax=40, mode="ragged")
    x = torch.from_numpy(g["logits"]).cuda()
    loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"], want_grad=False)
    want_loss, _, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"],
                                             precision="f64", want_grad=False)
    assert grad is None
    np.testing.assert_allclose(loss.cpu().numpy(), want_loss, rtol=LOSS_RTOL)


def test_random_shapes_sweep(common):
    """Seeded sweep over odd shapes (hypothesis-style, fixed seeds so the GPU box needs no database)."""
    rng = np.random.default_rng(2024)
    for i in range(12):
        C = int(rng.integers(2, 60))
        Lmax = int(rng.integers(1, 70))
        T = int(rng.integers(max(2 * Lmax + 2, 17), 2 * Lmax + 160))
        B = int(rng.integers(1, 7))
        g = make_batch(1000 + i, T=T, B=B, C=C, Lmax=Lmax, mode=["ragged", "full", "tight"][i % 3],
                       repeat_p=0.0 if C == 2 else 0.15, empty_row=bool(i % 2))
        want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
            g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
        loss, grad, status = _run_loss(common, g)
        _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status)


def _sparse_from_rows(rows, dtype):
    idx = np.asarray([(b, j) for b, r in enumerate(rows) for j in range(len(r))], np.int64).reshape(-1, 2)
    vals = np.asarray([v for r in rows for v in r], dtype)
    return idx, vals, np.asarray([len(rows), max(1, max(len(r) for r in rows))], np.int64)


def test_edit_distance_paths_match_oracle(common):
    """Bit-vector path (small symbol values, truth <= 1024), and the wavefront path it falls back to for
    long truths and for symbol values that do not fit the match-mask table."""
    from neuralasr_b200.networks.common import edit_distance
    rng = np.random.default_rng(5)
    cases = []
    for m, n, hi in [(1, 1, 3), (31, 40, 4), (32, 7, 4), (33, 90, 3), (64, 64, 2), (65, 300, 30), (200, 900, 37),
                     (1024, 700, 20), (1025, 300, 20), (1500, 1400, 5), (50, 60, 10 ** 6), (0, 5, 3), (7, 0, 3)]:
        cases.append((rng.integers(0, hi, m).tolist(), rng.integers(0, hi + 1, n).tolist()))
    truth_rows = [c[0] for c in cases]
    hyp_rows = [c[1] for c in cases]
    d, ler = edit_distance(_sparse_from_rows(hyp_rows, np.int64), _sparse_from_rows(truth_rows, np.int32))
    want = [o.levenshtein(h, t) for t, h in cases]
    assert d.cpu().numpy().tolist() == want


def test_edit_distance_lane_per_direction_kernel_matches_oracle(common, monkeypatch):
    """Truths of up to 224 symbols run in edit_distance_lanes_kernel (a lane per utterance and direction, 16
    utterances per warp, 2 / 4 / 7 words per lane); batches whose symbol values do not fit its table are handed to
    the warp-per-utterance kernel in the same call.  Both against the oracle, and against each other."""
    from neuralasr_b200.networks.common import edit_distance
    rng = np.random.default_rng(77)
    batches = []
    for mmax, hi in [(64, 5), (33, 37), (128, 37), (100, 3), (224, 37), (200, 60), (224, 2), (161, 500)]:
        for B in (1, 5, 16, 17, 40):
            truths, hyps = [], []
            for b in range(B):
                m = int(rng.integers(0, mmax + 1)) if b else mmax
                n = int(rng.choice([0, 1, 2, 3, 31, 32, 33, int(rng.integers(0, 700))]))
                if b == 2:
                    m = 0
                t = rng.integers(0, hi, m)
                if b % 3 == 0 and m and n:      # a hypothesis close to the truth: small distances
                    h = np.repeat(t, rng.integers(1, 3, m))[:n]
                    h = np.where(rng.random(h.size) < 0.1, rng.integers(0, hi + 1, h.size), h)
                else:
                    h = rng.integers(0, hi + 2, n)
                truths.append(t.tolist())
                hyps.append(h.tolist())
            batches.append((truths, hyps))
    # symbol values outside the table (declined as a whole warp), negative and huge hypothesis symbols (no match)
    t = rng.integers(0, 30, 150).tolist()
    batches.append(([t, [100000, 1, 2], t], [t[10:140], [1, 2, 100000], t[::-1]]))
    batches.append(([t, t[:5]], [[-3] + t[:100] + [2 ** 31 - 1] + t[100:], [4, 4, -1]]))

    def csr(rows):
        return (np.asarray([v for r in rows for v in r], np.int64),
                np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32))

    for truths, hyps in batches:
        want, want_ler = c_oracle.edit_distance(*csr(hyps), *csr(truths))
        hyp_sp, truth_sp = _sparse_from_rows(hyps, np.int64), _sparse_from_rows(truths, np.int32)
        d, ler = edit_distance(hyp_sp, truth_sp)
        assert d.cpu().numpy().tolist() == want.tolist()
        assert np.array_equal(ler.cpu().numpy(), want_ler)
        monkeypatch.setenv("NASR_LER_LANES", "0")
        d0, ler0 = edit_distance(hyp_sp, truth_sp)
        monkeypatch.delenv("NASR_LER_LANES")
        assert d0.cpu().numpy().tolist() == want.tolist()
        assert np.array_equal(ler0.cpu().numpy(), want_ler)


def test_batch_major_logits_without_transpose(common):
    """A [B,T,C] model output goes in as a strided [T,B,C] view (plain and DLPack entries): same loss, the
    gradient comes back in the batch-major layout, decode agrees."""
    g = make_batch(17, T=150, B=6, C=38, Lmax=30, mode="ragged")
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                                     g["seq_len"], precision="f64")
    x_btc = torch.from_numpy(np.ascontiguousarray(g["logits"].transpose(1, 0, 2))).cuda()     # [B,T,C]
    view = common.batch_major(x_btc)
    assert not view.is_contiguous()
    for use_dlpack in (False, True):
        loss, grad, status = common.ctc_loss_and_grad(view, _triple(g), g["seq_len"], use_dlpack=use_dlpack)
        assert grad.stride() == view.stride()
        np.testing.assert_allclose(loss.cpu().numpy(), want_loss, rtol=LOSS_RTOL, atol=1e-5)
        assert np.abs(grad.cpu().numpy() - want_grad).max() <= GRAD_ATOL
        assert np.abs(grad.transpose(0, 1).contiguous().cpu().numpy() - want_grad.transpose(1, 0, 2)).max() <= GRAD_ATOL
    xr = x_btc.clone().requires_grad_(True)
    common.loss(common.batch_major(xr), _triple(g), g["seq_len"]).backward()
    assert np.abs(xr.grad.cpu().numpy() - want_grad.transpose(1, 0, 2) / 6).max() <= GRAD_ATOL
    dec, _ = common.decoding(view, g["seq_len"])
    hv, _, _ = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    assert np.array_equal(dec.values.cpu().numpy(), hv)


def test_host_context_blocks_match_single_launch(common):
    """The HOST-buffer call splits a large batch into four utterance blocks (pitched copies, strided launches):
    results must equal the device path bit for bit."""
    from neuralasr_b200 import host
    g = make_batch(78, T=700, B=26, C=38, Lmax=60, mode="ragged")     # 2.8 MB... below the split threshold
    g2 = make_batch(79, T=900, B=70, C=38, Lmax=80, mode="ragged")    # 9.6 MB: split into 4 blocks of 17/18
    for gg in (g, g2):
        T, B, C = gg["logits"].shape
        ctx = host.HostContext(0, T, B, C, 80)
        out = ctx.step(gg["logits"], gg["label_values"], gg["label_offsets"], gg["seq_len"], want_hyp=True)
        loss, grad, status = _run_loss(common, gg)
        assert np.array_equal(out["loss"], loss) and np.array_equal(out["grad"], grad)
        assert np.array_equal(out["status"], status)
        hv, ho, _ = c_oracle.greedy_decode(gg["logits"], gg["seq_len"])
        want_d, want_ler = c_oracle.edit_distance(hv, ho, gg["label_values"], gg["label_offsets"])
        assert np.array_equal(out["hyp_len"], np.diff(ho)) and np.array_equal(out["dist"], want_d)
        assert np.array_equal(out["ler"], want_ler)
        for b in range(B):
            assert np.array_equal(out["hyp"][b, : out["hyp_len"][b]], hv[ho[b]:ho[b + 1]])
        ctx.close()


def test_step_functions_follow_the_reference_conventions(common):
    """train / validate / evaluate / decode (tfnetwork.py:166-190) on logits, dense padded labels in."""
    from neuralasr_b200.steps import CtcHead, dense_to_sparse
    g = make_batch(55, T=120, B=5, C=20, Lmax=12, mode="ragged", empty_row=False)
    lens = np.diff(g["label_offsets"])
    dense = np.zeros((5, lens.max()), np.int32)
    for b, lab in enumerate(g["labels_dense"]):
        dense[b, : len(lab)] = lab
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"],
                                                     g["seq_len"], precision="f64")
    hv, ho, _ = c_oracle.greedy_decode(g["logits"], g["seq_len"])
    _, want_ler = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    head = CtcHead()
    x = torch.from_numpy(g["logits"]).cuda().requires_grad_(True)
    loss_val, ler_val = head.train(x, dense, g["seq_len"], lens)
    assert abs(loss_val - want_loss.mean()) <= LOSS_RTOL * want_loss.mean()
    assert abs(ler_val - want_ler.astype(np.float64).mean()) < 1e-6
    assert np.abs(x.grad.cpu().numpy() - want_grad / 5).max() <= GRAD_ATOL and head.global_step == 1
    v = head.validate(x.detach(), dense, g["seq_len"], lens)
    assert isinstance(v, list) and abs(v[0] - loss_val) < 1e-6 and abs(v[1] - ler_val) < 1e-7
    values, l2, e2 = head.evaluate(x.detach(), dense, g["seq_len"], lens)
    assert np.array_equal(values, hv) and values.dtype == np.int64
    assert np.array_equal(head.decode(x.detach(), g["seq_len"]), hv)
    # batch-major head: same numbers from a [B,T,C] tensor
    bm = CtcHead(time_major=False)
    xb = torch.from_numpy(np.ascontiguousarray(g["logits"].transpose(1, 0, 2))).cuda()
    assert np.array_equal(bm.decode(xb, g["seq_len"]), hv)
    assert abs(bm.validate(xb, dense, g["seq_len"], lens)[0] - loss_val) < 1e-5
    # LAS-style metric caller: dense predictions and dense labels through dense_to_sparse (las.py:116-117)
    pred = np.zeros((5, 40), np.int64)
    for b in range(5):
        h = hv[ho[b]:ho[b + 1]][:40] + 1            # shift ids by one so that 0 can be the eos filler
        pred[b, : len(h)] = h
    lab1 = np.where(np.arange(dense.shape[1])[None, :] < lens[:, None], dense + 1, 0)
    m = common.label_error_rate(dense_to_sparse(pred), dense_to_sparse(lab1))
    want = [o.levenshtein((hv[ho[b]:ho[b + 1]][:40]).tolist(), g["labels_dense"][b].tolist()) for b in range(5)]
    assert m.distances.cpu().numpy().tolist() == want


def test_unaligned_wide_rows(common):
    """The wide-vocabulary variant moves rows with 16-byte accesses; a view whose rows are not 16-byte aligned must
    still be answered correctly (streamed rows peel to the boundary; a gradient buffer with another misalignment
    than the logits sends the call to the robust kernel), never faulted on."""
    g = make_batch(77, T=60, B=3, C=132, Lmax=12, mode="ragged")
    big = torch.zeros((60, 3, 135), device="cuda")
    view = big[:, :, 1:133]                       # class axis dense, rows start 4 bytes off a 16-byte boundary
    view.copy_(torch.from_numpy(g["logits"]))
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    loss, grad, status = common.ctc_loss_and_grad(view, _triple(g), g["seq_len"])
    _assert_loss_grad(loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy(), want_loss, want_grad, want_status)
    gbig = torch.zeros((60, 3, 135), device="cuda")
    gview = gbig[:, :, 2:134]                     # same strides as the logits view, another misalignment
    loss, grad, status = common.ctc_loss_and_grad(view, _triple(g), g["seq_len"], out_grad=gview)
    _assert_loss_grad(loss.cpu().numpy(), gview.cpu().numpy(), status.cpu().numpy(), want_loss, want_grad, want_status)


@pytest.mark.parametrize("C", [130, 132, 1024, 3001])
def test_greedy_decode_wide_rows_first_index_ties(common, C):
    """Wide rows go through the split arg-max kernel (16-byte loads when aligned, scalar otherwise): decoded ids,
    lengths and -sum(max logit) must equal the oracle's bit for bit, ties included (the earliest class wins)."""
    rng = np.random.default_rng(C)
    T, B = 70, 5
    x = rng.standard_normal((T, B, C)).astype(np.float32)
    x[::3, :, 7] = 9.0
    x[::3, :, C - 2] = 9.0            # two equal maxima in every third frame: class 7 must win
    x[5:9, 1, C - 1] = 12.0           # a run of blanks
    seq = np.array([T, 33, 1, 0, 64], np.int32)
    dec, nsl = common.decoding(torch.from_numpy(x).cuda(), seq)
    hv, ho, want_nsl = c_oracle.greedy_decode(x, seq)
    assert np.array_equal(dec.hyp_len.cpu().numpy(), np.diff(ho))
    assert np.array_equal(dec.values.cpu().numpy(), hv)
    assert np.array_equal(nsl.cpu().numpy().reshape(-1), want_nsl)


def test_loss_grad_and_decode_capture_into_a_cuda_graph(common):
    """include/nasr_ctc.h promises that the entry points only enqueue work on the given stream: a step (loss+grad,
    batch sums, greedy decode, beam search, label error rate) must be capturable into a CUDA graph and replay
    correctly on new logits."""
    g = make_batch(31, T=90, B=6, C=38, Lmax=20, mode="ragged")
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(g["logits"]).to(dev)
    lab = common.prepare_labels(_triple(g), dev)
    seq = torch.from_numpy(g["seq_len"]).to(dev)
    gbuf = torch.empty_like(x)

    def step():
        loss_b, grad, status = common.ctc_loss_and_grad(x, lab, seq, out_grad=gbuf)
        sums = common.batch_sums(loss_b=loss_b)
        dec, nsl = common.decoding(x, seq)
        dist, ler = common.edit_distance(dec, lab)
        bdec, blp = common.beam_decoding(x, seq, beam_width=16)
        return loss_b, status, sums, dec, dist, bdec, blp

    step()                                   # warm-up outside the capture: workspaces, function attributes
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_b, status, sums, dec, dist, bdec, blp = step()
    g2 = make_batch(32, T=90, B=6, C=38, Lmax=20, mode="ragged")
    x.copy_(torch.from_numpy(g2["logits"]))  # same labels and lengths, new logits
    graph.replay()
    torch.cuda.synchronize()
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g2["logits"], g["label_values"], g["label_offsets"], g["seq_len"],
                                                     precision="f64")
    np.testing.assert_allclose(loss_b.cpu().numpy(), want_loss, rtol=LOSS_RTOL)
    assert np.abs(gbuf.cpu().numpy() - want_grad).max() <= GRAD_ATOL
    assert abs(sums[0].item() - want_loss.sum()) <= LOSS_RTOL * want_loss.sum()
    hv, ho, _ = c_oracle.greedy_decode(g2["logits"], g["seq_len"])
    assert np.array_equal(dec.hyp_len.cpu().numpy(), np.diff(ho))
    want_d, _ = c_oracle.edit_distance(hv, ho, g["label_values"], g["label_offsets"])
    assert np.array_equal(dist.cpu().numpy(), want_d)
    hyp, hl, lp = c_oracle.beam_search(g2["logits"], g["seq_len"], 16, 1, True)
    assert np.array_equal(bdec[0].hyp_len.cpu().numpy(), hl[:, 0])
    assert np.allclose(blp.cpu().numpy()[:, 0], lp[:, 0], rtol=1e-6)


def test_wide_shapes_sweep(common):
    """Seeded sweep over wide vocabularies: register-held rows (C <= 1024, C % 4 == 0), streamed rows (wider, odd
    lengths), batch-major storage and rows offset by one float inside a wider buffer (misaligned)."""
    rng = np.random.default_rng(777)
    for case in range(36):
        C = int(rng.choice([65, 66, 131, 256, 1001, 1024, 1025, 2048, 3187, 6001, 8192]))
        Lmax = int(rng.integers(1, 120))
        T = int(rng.integers(max(2 * Lmax + 2, 17), 2 * Lmax + 100))
        B = int(rng.integers(1, 4))
        while T * B * C > 12_000_000 and B > 1:
            B -= 1
        if T * B * C > 12_000_000:
            Lmax = max(1, min(Lmax, 12_000_000 // (C * 4)))
            T = max(2 * Lmax + 2, 17)
        g = make_batch(5000 + case, T=T, B=B, C=C, Lmax=Lmax, mode=["ragged", "full", "tight"][case % 3],
                       peaky=bool(case % 2), empty_row=bool(case % 4 == 0))
        x = torch.from_numpy(g["logits"]).cuda()
        if case % 3 == 1:
            x = x.transpose(0, 1).contiguous().transpose(0, 1)
        elif case % 3 == 2:
            big = torch.zeros((T, B, C + 3), device="cuda")
            big[:, :, 1:C + 1] = x
            x = big[:, :, 1:C + 1]
        loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"])
        want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
            g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
        _assert_loss_grad(loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy(), want_loss, want_grad,
                          want_status)


def test_tf_published_basic_case_through_the_c_abi(common):
    """TensorFlow's ctc_loss_op_test.py basic case (tests/golden/tf_ctc_loss_op_test_basic.json): losses to the six
    published digits, the published gradient entries."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "tf_ctc_loss_op_test_basic.json")) as f:
        d = json.load(f)
    probs = np.stack([np.asarray(c["probs"], np.float64) for c in d["cases"]], axis=1)     # [T=5, B=2, C=6]
    labs = [np.asarray(c["labels"], np.int32) for c in d["cases"]]
    g = dict(logits=np.log(probs).astype(np.float32), label_values=np.concatenate(labs),
             label_offsets=np.asarray([0, labs[0].size, labs[0].size + labs[1].size], np.int32),
             seq_len=np.asarray([5, 5], np.int32))
    loss, grad, status = _run_loss(common, g)
    assert not status.any()
    for b, c in enumerate(d["cases"]):
        assert abs(loss[b] - c["loss"]) < 2e-5
        for t, k, want in c["grad_entries"]:
            assert abs(grad[t, b, k] - want) < 2e-6 + 1e-6


@pytest.mark.parametrize("blank", [0, 7])
def test_blank_anywhere_on_both_paths(common, debug_paths, blank):
    """The blank may be any class (the C-ABI takes it explicitly): same answers from the throughput kernel and from
    the robust one, labels may be every class but the blank."""
    g = make_batch(31, T=96, B=6, C=12, Lmax=14, mode="ragged")
    # make_batch draws labels from [0, C-2] with the blank last: move the classes around so that `blank` is the blank
    vals = g["label_values"].copy()
    vals[vals == blank] = 11
    g["label_values"] = vals
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64", blank=blank)
    assert not want_status.any()
    for path in (0, 1):
        debug_paths(path, 0)
        x = torch.from_numpy(g["logits"]).cuda()
        loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"], blank=blank)
        torch.cuda.synchronize()
        _assert_loss_grad(loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy(), want_loss, want_grad, want_status)


def test_full_size_cfg4_whole_batch_against_oracle(common):
    """BASELINE cfg4 at its own size (B=32, T=3000, L<=600): every utterance against the C oracle."""
    g = make_batch(404, T=3000, B=32, C=38, Lmax=600, mode="full", empty_row=False)
    want_loss, want_grad, want_status = c_oracle.ctc_loss_grad(
        g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    loss, grad, status = _run_loss(common, g)
    _assert_loss_grad(loss, grad, status, want_loss, want_grad, want_status)


def test_allreduce_scalars_through_the_c_abi(common, narrow_kernel):
    """nasr_allreduce_scalars on a one-rank NCCL communicator (the call a non-torch caller makes for the tower means,
    tfnetwork.py:135-136): the 4-vector comes back unchanged; NCCL is resolved from the process, not linked."""
    if narrow_kernel != "f64":
        pytest.skip("kernel-independent")
    import ctypes
    from neuralasr_b200 import _lib
    try:
        nccl = ctypes.CDLL("libnccl.so.2")
    except OSError:
        pytest.skip("libnccl.so.2 not loadable")
    comm = ctypes.c_void_p()
    devs = (ctypes.c_int * 1)(0)
    assert nccl.ncclCommInitAll(ctypes.byref(comm), 1, devs) == 0
    try:
        v = torch.tensor([1.5, 2.5, 3.0, 4.0], dtype=torch.float64, device="cuda:0")
        rc = _lib.load().nasr_allreduce_scalars(comm, ctypes.c_void_p(v.data_ptr()), 4,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert v.cpu().tolist() == [1.5, 2.5, 3.0, 4.0]
    finally:
        nccl.ncclCommDestroy(comm)


def test_device_side_label_ingest(common, narrow_kernel):
    """COO triple on the GPU -> CSR on the GPU: same offsets / values / longest row as the host
    conversion, a tower's row window re-based like tf.sparse_split, disorder flagged; and the loss takes it."""
    if narrow_kernel != "f64":
        pytest.skip("kernel-independent")
    import ctypes
    from neuralasr_b200 import _lib
    from neuralasr_b200.utils import sparse_to_csr, sparse_tuple_from, split_labels
    rng = np.random.default_rng(5)
    B, Lmax = 12, 9
    lens = rng.integers(0, Lmax + 1, size=B)
    lens[3] = 0
    lens[7] = Lmax
    dense = rng.integers(0, 30, size=(B, Lmax))
    idx, vals, shape = sparse_tuple_from(dense, lens)
    want_vals, want_offs, want_max = sparse_to_csr((idx, vals, shape))
    di, dv = torch.from_numpy(idx).cuda(), torch.from_numpy(vals.astype(np.int32)).cuda()
    lab = common.prepare_labels((di, dv, shape), torch.device("cuda", 0))
    torch.cuda.synchronize()
    assert np.array_equal(lab.offsets.cpu().numpy(), want_offs)
    assert np.array_equal(lab.values.cpu().numpy()[: want_vals.size], want_vals)
    info = lab.host_values.cpu().numpy()
    assert info[0] == 0 and info[1] == want_max == Lmax
    # a tower's block: rows [4, 8) of the same triple
    part = split_labels((idx, vals, shape), 3)[1]
    pv, po, _ = sparse_to_csr(part)
    lab2 = common.prepare_labels_device(di, dv, shape, row0=4, rows=4)
    torch.cuda.synchronize()
    assert np.array_equal(lab2.offsets.cpu().numpy(), po)
    assert np.array_equal(lab2.values.cpu().numpy()[: pv.size], pv)
    # rows out of order are flagged
    bad = idx.copy()
    bad[[0, -1]] = bad[[-1, 0]]
    with pytest.raises(ValueError):
        common.prepare_labels_device(torch.from_numpy(bad).cuda(), dv, shape, check=True)
    # and the loss accepts the device-resident triple
    g = make_batch(77, T=64, B=B, C=38, Lmax=Lmax, mode="full", empty_row=False)
    offs = g["label_offsets"]
    l2 = np.diff(offs)
    d2 = np.zeros((B, max(int(l2.max()), 1)), np.int32)
    for b in range(B):
        d2[b, : l2[b]] = g["label_values"][offs[b]:offs[b + 1]]
    i3, v3, s3 = sparse_tuple_from(d2, l2)
    x = torch.from_numpy(g["logits"]).cuda()
    loss_d, grad_d, _ = common.ctc_loss_and_grad(x, (torch.from_numpy(i3).cuda(), torch.from_numpy(v3.astype(np.int32)).cuda(), s3), g["seq_len"])
    loss_h, grad_h, _ = common.ctc_loss_and_grad(x, (i3, v3, s3), g["seq_len"])
    torch.cuda.synchronize()
    assert torch.equal(loss_d, loss_h) and torch.equal(grad_d, grad_h)


def test_tf_published_greedy_decoder_case_through_the_c_abi(common):
    """TensorFlow's own ctc_decoder_ops_test.py greedy case (tests/golden/tf_ctc_decoder_ops_test_greedy.json) through
    the C-ABI: the published SparseTensor triple and log probabilities, -inf logits included."""
    from test_oracle import tf_greedy_case
    g, x = tf_greedy_case()
    dec, nsl = common.decoding(torch.from_numpy(x).cuda(), g["seq_len"])
    idx, vals, shape = dec
    assert vals.tolist() == g["values"] and idx.tolist() == g["indices"] and shape.tolist() == g["dense_shape"]
    want_lp = np.array([np.sum(-np.log(np.asarray(p, np.float32))) for p in g["max_probs"]], np.float32)
    assert np.allclose(nsl[:, 0].cpu().numpy(), want_lp, rtol=1e-6)

"""GPU parity of the beam search decoder (csrc/ctc_beam.cu, through the C-ABI) against oracle/beam_oracle.py.

Label sequences must be identical and log probabilities equal to float32 rounding: both sides compute in
float64 with the same formulas and the same tie-break (score, kept prefix before new extension, prefix hash),
so they can only part on a score comparison closer than the two libms' last bit."""
import numpy as np
import pytest
import torch

from oracle import beam_oracle as bo

pytestmark = pytest.mark.gpu


def _run(x, seq_len, W=100, P=1, merge=True):
    from neuralasr_b200.networks import common
    dec, lp = common.beam_decoding(torch.from_numpy(x).cuda(), seq_len, beam_width=W, top_paths=P,
                                   merge_repeated=merge)
    torch.cuda.synchronize()
    out = []
    for p in range(P):
        hyp = dec[p].hyp.cpu().numpy()
        hl = dec[p].hyp_len.cpu().numpy()
        out.append([hyp[b, :hl[b]].tolist() for b in range(x.shape[1])])
    return out, lp.cpu().numpy()


def _check(x, seq_len, W=100, P=1, merge=True):
    got, lp = _run(x, seq_len, W, P, merge)
    for b in range(x.shape[1]):
        want = bo.beam_search_one(x[: int(seq_len[b]), b, :].astype(np.float64), W, merge, top_paths=P)
        for p in range(P):
            if p < len(want):
                assert got[p][b] == want[p][0], (b, p)
                assert abs(lp[b, p] - want[p][1]) <= 1e-6 * max(1.0, abs(want[p][1])), (b, p, lp[b, p], want[p][1])
            else:
                assert got[p][b] == [] and lp[b, p] == -np.inf


def _peaky(rng, T, B, C, noise=1.0):
    """Planted alignment: each label held 1-4 frames with blanks between, +8 on the planted class."""
    x = rng.normal(size=(T, B, C)).astype(np.float32) * noise
    for b in range(B):
        t = 0
        while t < T:
            c = rng.integers(0, C - 1) if rng.random() < 0.6 else C - 1
            hold = int(rng.integers(1, 5))
            x[t:t + hold, b, c] += 8.0
            t += hold
    return x


@pytest.mark.parametrize("T,B,C,W", [(12, 3, 5, 4), (30, 4, 38, 100), (25, 2, 38, 16), (40, 2, 7, 128), (9, 2, 3, 100)])
def test_random_logits(T, B, C, W):
    rng = np.random.default_rng(T * 1000 + C)
    x = (rng.normal(size=(T, B, C)) * 3).astype(np.float32)
    seq = np.array([T] + list(rng.integers(1, T + 1, size=B - 1)), np.int32)
    _check(x, seq, W)


def test_peaky_logits_c38_w100():
    rng = np.random.default_rng(7)
    x = _peaky(rng, 60, 3, 38)
    _check(x, np.array([60, 45, 33], np.int32), 100)


def test_top_paths_and_no_merge():
    rng = np.random.default_rng(11)
    x = (rng.normal(size=(20, 2, 6)) * 2).astype(np.float32)
    _check(x, np.array([20, 13], np.int32), W=32, P=5, merge=True)
    _check(x, np.array([20, 13], np.int32), W=32, P=3, merge=False)


def test_exact_ties_uniform_logits():
    # every extension of a frame ties exactly: the (kept first, hash) order decides, identically on both sides
    x = np.zeros((8, 2, 6), np.float32)
    _check(x, np.array([8, 5], np.int32), W=10, P=4, merge=False)


def test_merge_repeated_quirk_and_empty():
    a, b, blank = 0, 1, 2
    x = np.full((6, 3, 3), -4.0, np.float32)
    for t, c in enumerate([a, a, blank, a, b, b]):
        x[t, :, c] = 4.0
    got, lp = _run(x, np.array([6, 3, 0], np.int32), 100, 1, True)
    assert got[0][0] == [a, b] and got[0][1] == [a] and got[0][2] == []
    assert lp[2, 0] == 0.0
    got, _ = _run(x, np.array([6, 3, 0], np.int32), 100, 1, False)
    assert got[0][0] == [a, a, b]


def test_wide_vocabulary_and_batch_major_view():
    rng = np.random.default_rng(3)
    xb = _peaky(rng, 24, 2, 300).transpose(1, 0, 2).copy()      # [B, T, C] as a model produces it
    from neuralasr_b200.networks import common
    view = common.batch_major(torch.from_numpy(xb).cuda())      # [T, B, C] strided view, no copy
    dec, lp = common.beam_decoding(view, np.array([24, 17], np.int32), beam_width=20)
    hyp, hl = dec[0].hyp.cpu().numpy(), dec[0].hyp_len.cpu().numpy()
    for b, Tb in enumerate([24, 17]):
        want = bo.beam_search_one(xb[b, :Tb].astype(np.float64), 20, True)
        assert hyp[b, :hl[b]].tolist() == want[0]
        assert abs(lp[b, 0].item() - want[1]) <= 1e-6 * max(1.0, abs(want[1]))


def test_beam_feeds_label_error_rate_like_create_model_create_metric():
    from neuralasr_b200.networks import common
    from neuralasr_b200.utils import sparse_tuple_from
    rng = np.random.default_rng(5)
    x = _peaky(rng, 50, 4, 38, noise=0.5)
    seq = np.array([50, 50, 40, 30], np.int32)
    model, log_prob = common.create_model_beam(torch.from_numpy(x).cuda(), seq)
    truth = [bo.beam_search_one(x[: seq[b], b].astype(np.float64))[0] or [0] for b in range(4)]
    truth[1] = truth[1][:-1] + [(truth[1][-1] + 1) % 37]          # one substitution
    labels = sparse_tuple_from([np.array(t) for t in truth], [len(t) for t in truth])
    ler = common.label_error_rate(model, labels)
    assert ler.distances.cpu().tolist()[1] == 1
    assert log_prob.shape == (4, 1)
    assert tuple(model.dense_shape.cpu().tolist())[0] == 4


def test_unsupported_and_bad_arguments():
    from neuralasr_b200.networks import common
    x = torch.zeros((4, 1, 5), device="cuda")
    with pytest.raises(ValueError):
        common.beam_decoding(x, np.array([4], np.int32), beam_width=4, top_paths=8)
    with pytest.raises(ValueError):
        common.beam_decoding(x.cpu(), np.array([4], np.int32))


def _check_c(x, seq, W=100, P=1, merge=True, allow=0):
    """Against the C port (oracle/beam_oracle.c), which is fast enough for full-length utterances."""
    from oracle import c_oracle
    got, lp = _run(x, seq, W, P, merge)
    hyp, hl, want_lp = c_oracle.beam_search(x, seq, W, P, merge)
    bad = 0
    for b in range(x.shape[1]):
        for p in range(P):
            same = got[p][b] == hyp[b, p, : hl[b, p]].tolist() and \
                abs(lp[b, p] - want_lp[b, p]) <= 1e-6 * max(1.0, abs(want_lp[b, p]))
            bad += not same
    assert bad <= allow, "%d of %d paths differ" % (bad, x.shape[1] * P)


def test_full_length_utterances_against_c_port():
    # BASELINE cfg2 shape per utterance (T=800, C=38, beam 100), random and peaky, ragged lengths
    rng = np.random.default_rng(42)
    x = (rng.normal(size=(800, 6, 38)) * 3).astype(np.float32)
    _check_c(x, np.array([800, 800, 641, 400, 17, 0], np.int32))
    _check_c(_peaky(rng, 800, 6, 38), np.array([800, 799, 555, 300, 1, 64], np.int32))


def test_cfg1_shape_and_reference_config_vocabulary():
    rng = np.random.default_rng(43)
    _check_c(_peaky(rng, 500, 16, 38), np.full(16, 500, np.int32))          # cfg1: B=16, T=500, C=38
    _check_c(_peaky(rng, 300, 4, 41), np.full(4, 300, np.int32))            # C=41: 8000sr config with ^ and $
    _check_c((rng.normal(size=(120, 3, 41)) * 3).astype(np.float32), np.full(3, 120, np.int32))  # > 4096 candidates


def test_wide_vocabulary_against_c_port():
    rng = np.random.default_rng(44)
    _check_c(_peaky(rng, 200, 3, 1024), np.array([200, 150, 77], np.int32))
    _check_c((rng.normal(size=(40, 2, 1024)) * 3).astype(np.float32), np.array([40, 33], np.int32), W=64, P=2)


def test_widest_vocabulary():
    rng = np.random.default_rng(48)
    _check_c(_peaky(rng, 40, 2, 8192), np.array([40, 23], np.int32), W=100)
    _check_c((rng.normal(size=(12, 2, 5000)) * 3).astype(np.float32), np.array([12, 7], np.int32), W=64, P=2)
    from neuralasr_b200.networks import common
    with pytest.raises(Exception):
        common.beam_decoding(torch.zeros((4, 1, 8193), device="cuda"), np.array([4], np.int32))


def test_wide_beam():
    rng = np.random.default_rng(45)
    _check_c((rng.normal(size=(60, 2, 38)) * 2).astype(np.float32), np.array([60, 41], np.int32), W=512, P=3)


def test_step_functions_with_the_beam_decoder():
    """evaluate()/decode() of the step-function mirror (tfnetwork.py:172-181) with create_model's real decoder."""
    from neuralasr_b200.steps import CtcHead
    from oracle import c_oracle
    rng = np.random.default_rng(46)
    T, B, C, L = 80, 4, 38, 12
    x = _peaky(rng, T, B, C, noise=0.5)
    seq = np.array([80, 80, 61, 47], np.int32)
    dense = rng.integers(1, C - 1, size=(B, L))
    lens = np.array([12, 7, 9, 3])
    head = CtcHead(decoder="beam")
    values, mean_loss, mean_ler = head.evaluate(torch.from_numpy(x).cuda(), dense, seq, lens)
    hyp, hl, _ = c_oracle.beam_search(x, seq, 100, 1, True)
    want = np.concatenate([hyp[b, 0, : hl[b, 0]] for b in range(B)])
    assert np.array_equal(values, want)
    assert np.array_equal(head.decode(torch.from_numpy(x).cuda(), seq), want)
    offs = np.concatenate([[0], np.cumsum(hl[:, 0])]).astype(np.int32)
    tv = np.concatenate([dense[b, : lens[b]] for b in range(B)]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    _, want_ler = c_oracle.edit_distance(want, offs, tv, to)
    assert abs(float(mean_ler) - float(want_ler.astype(np.float64).mean())) < 1e-6
    assert np.isfinite(mean_loss)
    with pytest.raises(ValueError):
        CtcHead(decoder="viterbi")


def test_full_size_properties_cfg3():
    """BASELINE cfg3 (B=256, T=1000, C=38, width 100): size-independent properties instead of an oracle run.
    With a sharply planted alignment the most probable labelling is the collapsed arg-max path, so the beam's top
    path (repeats NOT merged in the output) must equal the greedy decoder's; the search is deterministic; log
    probabilities are <= 0 and ordered."""
    from neuralasr_b200.networks import common
    g = torch.Generator(device="cuda").manual_seed(5)
    T, B, C = 1000, 256, 38
    x = torch.randn((T, B, C), device="cuda", generator=g)
    cls = torch.randint(0, C, (T // 2, B), device="cuda", generator=g).repeat_interleave(2, 0)
    x.scatter_add_(2, cls.unsqueeze(-1), torch.full((T, B, 1), 30.0, device="cuda"))
    seq = torch.full((B,), T, dtype=torch.int32)
    seq[1], seq[2] = 517, 1
    dec, lp = common.beam_decoding(x, seq, beam_width=100, top_paths=2, merge_repeated=False)
    greedy, _ = common.decoding(x, seq)
    assert torch.equal(dec[0].hyp_len, greedy.hyp_len)
    keep = torch.arange(T, device="cuda")[None, :] < greedy.hyp_len[:, None]
    assert torch.equal(dec[0].hyp[keep], greedy.hyp[keep])
    assert (lp <= 0).all() and (lp[:, 0] >= lp[:, 1]).all() and (lp[:, 0] > -1e-3).all()
    dec2, lp2 = common.beam_decoding(x, seq, beam_width=100, top_paths=2, merge_repeated=False)
    assert torch.equal(lp, lp2) and torch.equal(dec[1].hyp_len, dec2[1].hyp_len)
    assert torch.equal(dec[0].hyp[keep], dec2[0].hyp[keep])


def test_host_buffer_step_with_the_beam_decoder():
    """nasr_host_ctc_step after nasr_host_ctx_set_decoder(beam): numpy in, what train() fetches out."""
    from neuralasr_b200 import host
    from oracle import c_oracle
    rng = np.random.default_rng(47)
    T, B, C, L = 120, 24, 38, 10
    x = _peaky(rng, T, B, C, noise=0.7)
    seq = rng.integers(60, T + 1, size=B).astype(np.int32)
    lens = rng.integers(1, L + 1, size=B)
    vals = rng.integers(0, C - 1, size=int(lens.sum())).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ctx = host.HostContext(0, T, B, C, L, decoder="beam", beam_width=50)
    out = ctx.step(x, vals, offs, seq, want_hyp=True)
    hyp, hl, lp = c_oracle.beam_search(x, seq, 50, 1, True)
    assert np.array_equal(out["hyp_len"], hl[:, 0])
    for b in range(B):
        assert np.array_equal(out["hyp"][b, : hl[b, 0]], hyp[b, 0, : hl[b, 0]])
    assert np.allclose(out["neg_sum_logits"], lp[:, 0], rtol=1e-6)
    hv = np.concatenate([hyp[b, 0, : hl[b, 0]] for b in range(B)])
    ho = np.concatenate([[0], np.cumsum(hl[:, 0])]).astype(np.int32)
    want_d, want_ler = c_oracle.edit_distance(hv, ho, vals, offs)
    assert np.array_equal(out["dist"], want_d) and np.allclose(out["ler"], want_ler)
    want_loss, _, _ = c_oracle.ctc_loss_grad(x, vals, offs, seq, precision="f64", want_grad=False)
    assert np.allclose(out["loss"], want_loss, rtol=1e-4)
    ctx.set_decoder("greedy")
    g = ctx.step(x, vals, offs, seq)
    gv, go, _ = c_oracle.greedy_decode(x, seq)
    assert np.array_equal(g["hyp_len"], np.diff(go))
    ctx.close()


def test_kernel_reproduces_the_golden_fixture():
    """tests/golden/beam_bruteforce.npz: exact labelling probabilities by enumeration of every alignment.  Width 512
    is at least the number of prefixes of these utterances, so the kernel's search is exhaustive here."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "beam_bruteforce.npz"))
    checked = 0
    for i in range(int(g["cases"])):
        x, n = g["x%d" % i], int(g["n%d" % i])
        T, C = x.shape
        assert sum((C - 1) ** k for k in range(T + 1)) <= 512
        got, lp = _run(x[:, None, :].copy(), np.array([T], np.int32), W=512, P=max(n, 1), merge=False)
        for j in range(n):
            want = g["labels%d" % i][j, : g["lens%d" % i][j]].tolist()
            assert got[j][0] == want, (i, j)
            assert abs(lp[0, j] - g["logp%d" % i][j]) <= 1e-6 * max(1.0, abs(g["logp%d" % i][j]))
            checked += 1
    assert checked >= 40


@pytest.mark.parametrize("W", [1, 2, 3])
def test_narrowest_beams(W):
    rng = np.random.default_rng(50 + W)
    x = (rng.normal(size=(30, 3, 6)) * 1.5).astype(np.float32)
    _check_c(x, np.array([30, 21, 9], np.int32), W=W, P=W, merge=False)


def test_blank_elsewhere_than_last_class():
    from neuralasr_b200.networks import common
    rng = np.random.default_rng(53)
    x = (rng.normal(size=(25, 2, 7)) * 2).astype(np.float32)
    seq = np.array([25, 14], np.int32)
    for blank in (0, 3):
        dec, lp = common.beam_decoding(torch.from_numpy(x).cuda(), seq, beam_width=20, blank=blank)
        hyp, hl = dec[0].hyp.cpu().numpy(), dec[0].hyp_len.cpu().numpy()
        for b in range(2):
            want = bo.beam_search_one(x[: seq[b], b].astype(np.float64), 20, True, blank=blank)
            assert hyp[b, : hl[b]].tolist() == want[0]
            assert abs(lp[b, 0].item() - want[1]) <= 1e-6 * max(1.0, abs(want[1]))


def test_wide_vocabulary_ties_through_the_second_stage_bound():
    # C > 64 runs ctc_beam_kernel<true>: exact ties (uniform rows, and rows with a few equal peaks) must come out as
    # in the oracle whatever the second-stage bound prunes
    x = np.zeros((6, 2, 100), np.float32)
    _check_c(x, np.array([6, 4], np.int32), W=10, P=4, merge=False)
    x[:, :, [3, 17, 64]] = 5.0
    x[2:, 0, 99] = 5.0
    _check_c(x, np.array([6, 6], np.int32), W=10, P=4, merge=False)
    rng = np.random.default_rng(54)
    _check_c((rng.normal(size=(50, 3, 65)) * 3).astype(np.float32), np.array([50, 50, 31], np.int32), W=100, P=2)


def test_randomised_shapes_against_the_c_port():
    """Seeded sweep over shapes, widths, blank positions and four kinds of logits.  The C port reports the smallest
    difference of totals over all decisions it took for an utterance; where that margin is within rounding of zero
    (quantised logits make different prefixes mathematically equal) the outcome hangs on the last bit of exp/log and
    the comparison is skipped — everywhere else the paths must be identical."""
    from neuralasr_b200.networks import common
    from oracle import c_oracle
    rng = np.random.default_rng(20240)
    compared = skipped = 0
    for case in range(240):
        C = int(rng.choice([2, 3, 5, 12, 38, 41, 64, 65, 100, 300, 1024]))
        T = int(rng.integers(1, 70))
        B = int(rng.integers(1, 5))
        W = int(rng.choice([1, 2, 3, 7, 16, 100, 128, 300]))
        P = int(min(W, rng.integers(1, 4)))
        merge = bool(rng.integers(0, 2))
        kind = case % 4
        if kind == 0:
            x = (rng.normal(size=(T, B, C)) * rng.choice([0.3, 1.0, 3.0, 8.0])).astype(np.float32)
        elif kind == 1:                                   # planted alignment
            x = rng.normal(size=(T, B, C)).astype(np.float32)
            np.put_along_axis(x, rng.integers(0, C, size=(T, B))[:, :, None], 8.0, axis=2)
        elif kind == 2:                                   # quantised logits: many exact and near ties
            x = rng.integers(-2, 3, size=(T, B, C)).astype(np.float32)
        else:                                             # extreme range
            x = (rng.normal(size=(T, B, C)) * 40).astype(np.float32)
        seq = rng.integers(0, T + 1, size=B).astype(np.int32)
        seq[0] = T
        blank = C - 1 if case % 5 else int(rng.integers(0, C))
        dec, lp = common.beam_decoding(torch.from_numpy(x).cuda(), seq, beam_width=W, top_paths=P,
                                       merge_repeated=merge, blank=blank)
        hyp, hl, want, margin = c_oracle.beam_search(x, seq, W, P, merge, blank=blank, with_margin=True)
        lp = lp.cpu().numpy()
        for b in range(B):
            # quantised logits: bitwise-equal totals of different prefixes may be one ulp apart under another libm
            if margin[b, 0] < 1e-9 or (kind == 2 and margin[b, 1] > 0):
                skipped += 1
                continue
            compared += 1
            for p in range(P):
                gh, gl = dec[p].hyp[b].cpu().numpy(), int(dec[p].hyp_len[b])
                assert gl == hl[b, p] and np.array_equal(gh[:gl], hyp[b, p, : hl[b, p]]), (case, kind, b, p)
                if np.isfinite(want[b, p]):
                    assert abs(lp[b, p] - want[b, p]) <= 1e-6 * max(1.0, abs(want[b, p])), (case, kind, b, p)
                else:
                    assert lp[b, p] == want[b, p]
    assert compared >= 350 and skipped <= compared // 2, (compared, skipped)


def test_full_size_cfg3_whole_batch_against_the_c_port():
    """BASELINE cfg3 (B=256, T=1000, C=38, width 100) on the bench's N(0,1)*3 logits: every utterance's top path and
    log probability against the C port (a second and a half on the box's host threads)."""
    from oracle import c_oracle
    rng = np.random.default_rng(1234)
    T, B, C = 1000, 256, 38
    x = rng.standard_normal((T, B, C), dtype=np.float32) * np.float32(3.0)
    seq = np.full(B, T, np.int32)
    seq[3], seq[200] = 611, 17
    got, lp = _run(x, seq, 100, 1, True)
    hyp, hl, want, margin = c_oracle.beam_search(x, seq, 100, 1, True, with_margin=True)
    # continuous logits: decisions within rounding are rare; bitwise-equal totals (margin[:, 1]) come from twins --
    # two labels with equal float32 logits in a frame -- which every implementation ties and the hash decides
    assert (margin[:, 0] > 1e-9).sum() >= B - 8
    for b in range(B):
        if margin[b, 0] <= 1e-9:
            continue
        assert got[0][b] == hyp[b, 0, : hl[b, 0]].tolist(), b
        assert abs(lp[b, 0] - want[b, 0]) <= 1e-6 * abs(want[b, 0]), b


def test_zero_frames():
    from neuralasr_b200.networks import common
    x = torch.zeros((0, 3, 7), device="cuda")
    dec, lp = common.beam_decoding(x, np.zeros(3, np.int32), beam_width=8, top_paths=2)
    assert dec[0].hyp_len.cpu().tolist() == [0, 0, 0] and dec[1].hyp_len.cpu().tolist() == [0, 0, 0]
    assert lp[:, 0].cpu().tolist() == [0.0, 0.0, 0.0] and torch.isinf(lp[:, 1]).all()


def test_tf_published_beam_search_case_through_the_c_abi():
    """TensorFlow's own ctc_decoder_ops_test.py beam search case (tests/golden/tf_ctc_decoder_ops_test_beam.json):
    the kernel, through the C-ABI, returns the published hypotheses [1, 0] and [1] with log probabilities -5.811451
    and -6.63339 (float32 tolerance 5e-6); the two padding frames beyond seq_len are ignored as TF ignores them."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "tf_ctc_decoder_ops_test_beam.json")) as f:
        g = json.load(f)
    p = np.asarray(g["input_prob_matrix"], np.float32) + np.float32(g["offset"])
    x = np.ascontiguousarray(np.concatenate([p, np.zeros((g["padding_frames"], p.shape[1]), np.float32)])[:, None, :])
    # a batch of three copies: the case must not depend on the batch position
    x = np.ascontiguousarray(np.repeat(x, 3, axis=1))
    got, lp = _run(x, np.full(3, g["seq_len"], np.int32), W=g["beam_width"], P=g["top_paths"],
                   merge=g["merge_repeated"])
    for b in range(3):
        for j, (lab, v) in enumerate(zip(g["decoded"], g["log_prob"])):
            assert got[j][b] == lab, (b, j, got[j][b])
            assert abs(lp[b, j] - v) < 5e-6, (b, j, lp[b, j])

"""GPU parity of the model tails' affine projection (csrc/affine.cu, through the C-ABI) against oracle/affine_oracle.py
(float64 numpy restatement of networks/bilstm_ctc_net.py:31-48).  The kernels multiply in 3xTF32 with float32
accumulation: the bound below is 4e-6 * (|H| . |W|) per element (a float32 matmul's own rounding is ~1e-6 of it)."""
import numpy as np
import pytest
import torch

from oracle import affine_oracle as ao

pytestmark = pytest.mark.gpu

REL = 4e-6


def _close(got, want, bound):
    err = np.abs(got.astype(np.float64) - want)
    assert np.all(err <= REL * bound + 1e-30), float((err / (bound + 1e-30)).max())


def _case(seed, rows, K, C, scale=1.0):
    rng = np.random.default_rng(seed)
    H = (rng.standard_normal((rows, K)) * scale).astype(np.float32)
    W = (rng.standard_normal((K, C)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32)
    return H, W, b


@pytest.fixture(params=["tcgen05", "mma.sync"])
def forward_path(request, monkeypatch):
    """The forward and dH have two kernels each: the tcgen05 / TMEM ones (csrc/affine_tc.cu: C <= 40, 16-byte aligned
    rows, at least 128 rows and 32 k; csrc/affine_tc_dh.cu: C <= 40, contiguous dlogits, at least 128 rows) and the
    mma.sync ones that take everything else.  NASR_AFFINE_TC=0 forces the second."""
    monkeypatch.setenv("NASR_AFFINE_TC", "1" if request.param == "tcgen05" else "0")
    return request.param


@pytest.mark.parametrize("rows,K,C", [(1, 1, 1), (16, 8, 8), (37, 500, 38), (1000, 500, 38), (4099, 500, 38),
                                        (130, 13, 5), (257, 100, 41), (300, 700, 38), (64, 1500, 7),
                                        (500, 500, 1024), (333, 250, 129), (64, 32, 40), (129, 257, 1),
                                        (20000, 500, 38), (128, 32, 40), (256, 260, 3), (300, 96, 40)])
def test_forward_matches_oracle(rows, K, C, forward_path):
    from neuralasr_b200.networks import common
    H, W, b = _case(rows * 7 + K + C, rows, K, C)
    got = common.affine_logits(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    _close(got.cpu().numpy(), ao.affine_logits(H, W, b), np.abs(H).astype(np.float64) @ np.abs(W) + np.abs(b))
    got = common.affine_logits(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda())
    _close(got.cpu().numpy(), ao.affine_logits(H, W), np.abs(H).astype(np.float64) @ np.abs(W))


@pytest.mark.parametrize("rows,K,C", [(1, 1, 1), (16, 8, 8), (37, 500, 38), (1000, 500, 38), (4099, 500, 38),
                                        (130, 13, 5), (257, 100, 41), (300, 700, 38), (500, 500, 1024),
                                        (20000, 500, 38), (128, 8, 40), (999, 129, 7), (640, 256, 38)])
def test_backward_matches_oracle(rows, K, C, forward_path):
    from neuralasr_b200.networks import common
    H, W, _ = _case(rows * 3 + K + C, rows, K, C)
    rng = np.random.default_rng(5)
    dL = rng.standard_normal((rows, C)).astype(np.float32)
    dH, dW, db = common.affine_backward(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda(),
                                        torch.from_numpy(dL).cuda())
    torch.cuda.synchronize()
    wH, wW, wb = ao.affine_backward(H, W, dL)
    _close(dH.cpu().numpy(), wH, np.abs(dL).astype(np.float64) @ np.abs(W).T)
    _close(dW.cpu().numpy(), wW, np.abs(H).astype(np.float64).T @ np.abs(dL))
    # db is a plain float32 sum over the rows (per-CTA partials, then a fixed-order reduction)
    assert np.all(np.abs(db.cpu().numpy() - wb) <= 2e-6 * np.abs(dL).astype(np.float64).sum(0) + 1e-30)
    # each output alone
    only_dH, none_w, none_b = common.affine_backward(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda(),
                                                     torch.from_numpy(dL).cuda(), True, False, False)
    assert none_w is None and none_b is None and torch.equal(only_dH, dH)
    _, only_dW, _ = common.affine_backward(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda(),
                                           torch.from_numpy(dL).cuda(), False, True, False)
    assert torch.equal(only_dW, dW)          # deterministic: per-CTA partials added in a fixed order


def test_strided_and_misaligned_rows(forward_path):
    from neuralasr_b200.networks import common
    H, W, b = _case(11, 200, 500, 38)
    big = torch.zeros((200, 503), dtype=torch.float32, device="cuda")
    view = big[:, 1:501]                                  # rows start 4 bytes off a 16-byte boundary, stride 503
    view.copy_(torch.from_numpy(H))
    got = common.affine_logits(view, torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda())
    _close(got.cpu().numpy(), ao.affine_logits(H, W, b), np.abs(H).astype(np.float64) @ np.abs(W) + np.abs(b))
    wide = torch.zeros((200, 504), dtype=torch.float32, device="cuda")
    aligned = wide[:, 4:504]                              # 16-byte aligned rows at stride 504: a strided tensor map
    aligned.copy_(torch.from_numpy(H))
    got = common.affine_logits(aligned, torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda())
    _close(got.cpu().numpy(), ao.affine_logits(H, W, b), np.abs(H).astype(np.float64) @ np.abs(W) + np.abs(b))
    out = torch.zeros((200, 39), dtype=torch.float32, device="cuda")[:, :38]   # odd row stride: scalar stores
    common.affine_logits(view, torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda(), out=out)
    _close(out.cpu().numpy(), ao.affine_logits(H, W, b), np.abs(H).astype(np.float64) @ np.abs(W) + np.abs(b))
    dL = torch.randn((200, 38), device="cuda")
    dH, dW, db = common.affine_backward(view, torch.from_numpy(W).cuda(), dL)
    wH, wW, wb = ao.affine_backward(H, W, dL.cpu().numpy())
    _close(dW.cpu().numpy(), wW, np.abs(H).astype(np.float64).T @ np.abs(dL.cpu().numpy()))
    _close(dH.cpu().numpy(), wH, np.abs(dL.cpu().numpy()).astype(np.float64) @ np.abs(W).T)


def test_tail_feeds_the_loss_without_a_transpose_and_backpropagates():
    """bilstm_ctc_net.py:31-52 end to end: projection -> (no transpose) -> create_loss, and the gradients of the mean
    loss w.r.t. outputs, W and b, against the oracle chain (float64 projection, C oracle for the loss)."""
    from neuralasr_b200.networks import common
    from neuralasr_b200.utils import sparse_tuple_from
    from oracle import c_oracle
    rng = np.random.default_rng(3)
    B, T, K, C, Lmax = 6, 70, 500, 38, 12
    outputs = rng.standard_normal((B, T, K)).astype(np.float32)
    W = (rng.standard_normal((K, C)) * (3.0 / np.sqrt(K))).astype(np.float32)
    b = (rng.standard_normal(C) * 0.1).astype(np.float32)
    lens = rng.integers(1, Lmax + 1, size=B)
    dense = rng.integers(0, C - 1, size=(B, Lmax))
    labels = sparse_tuple_from(dense, lens)
    seq = np.full(B, T, np.int32)
    seq[2] = 50
    o_t = torch.from_numpy(outputs).cuda().requires_grad_(True)
    W_t = torch.from_numpy(W).cuda().requires_grad_(True)
    b_t = torch.from_numpy(b).cuda().requires_grad_(True)
    logits = common.affine_projection(o_t, W_t, b_t, B)
    assert tuple(logits.shape) == (T, B, C) and logits.stride(0) == C and logits.stride(1) == T * C
    mean = common.loss(logits, labels, seq)
    mean.backward()
    torch.cuda.synchronize()
    want_logits = ao.tail(outputs, W, b, B).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(want_logits, labels[1], offs, seq, precision="f64")
    assert np.allclose(mean.per_utterance.cpu().numpy(), want_loss, rtol=1e-4)
    dL = (want_grad / B).transpose(1, 0, 2).reshape(B * T, C)           # batch-major rows, as the projection wrote them
    wH, wW, wb = ao.affine_backward(outputs.reshape(-1, K), W, dL)
    assert np.abs(o_t.grad.cpu().numpy().reshape(-1, K) - wH).max() < 1e-4
    assert np.abs(W_t.grad.cpu().numpy() - wW).max() < 1e-4 * max(1.0, np.abs(wW).max())
    assert np.abs(b_t.grad.cpu().numpy() - wb).max() < 1e-4
    # the BiLSTM tail's reshape of the (forward, backward) pair: rows stack both directions -> [B, 2T, C]
    pair = torch.stack([o_t.detach(), o_t.detach().flip(1)])
    both = common.affine_projection(pair, W_t.detach(), b_t.detach(), B)
    assert tuple(both.shape) == (2 * T, B, C)
    want = ao.tail(pair.cpu().numpy(), W, b, B)
    assert np.abs(both.cpu().numpy() - want).max() < 1e-4


def test_full_size_linearity_cfg3_rows(forward_path):
    """BASELINE cfg3's row count (B*T = 256000, K = 500, C = 38), size-independent properties: linearity in H,
    the bias reaching every row, 4096 sampled rows against the oracle, and dW against the oracle on the whole batch."""
    from neuralasr_b200.networks import common
    rows, K, C = 256000, 500, 38
    g = torch.Generator(device="cuda").manual_seed(1)
    H = torch.randn((rows, K), device="cuda", generator=g)
    W = torch.randn((K, C), device="cuda", generator=g) / K ** 0.5
    b = torch.randn((C,), device="cuda", generator=g)
    y = common.affine_logits(H, W, b)
    y2 = common.affine_logits(H * 2.0, W, b)      # scaling by a power of two is exact in every split
    assert torch.equal(y2 - b, (y - b) * 2.0) or float(((y2 - b) - (y - b) * 2.0).abs().max()) < 1e-5
    idx = torch.randint(0, rows, (4096,), device="cuda", generator=g)
    Hs, Wn, bn = H[idx].cpu().numpy(), W.cpu().numpy(), b.cpu().numpy()
    _close(y[idx].cpu().numpy(), ao.affine_logits(Hs, Wn, bn), np.abs(Hs).astype(np.float64) @ np.abs(Wn) + np.abs(bn))
    dL = torch.randn((rows, C), device="cuda", generator=g) * 1e-3
    dH, dW, db = common.affine_backward(H, W, dL)
    ref_dW = (H.double().T @ dL.double()).cpu().numpy()
    bound = (H.double().abs().T @ dL.double().abs()).cpu().numpy()
    _close(dW.cpu().numpy(), ref_dW, bound)
    assert np.all(np.abs(db.cpu().numpy() - dL.double().sum(0).cpu().numpy()) <= 2e-6 * dL.double().abs().sum(0).cpu().numpy())
    dLs = dL[idx].cpu().numpy()
    _close(dH[idx].cpu().numpy(), dLs.astype(np.float64) @ Wn.T.astype(np.float64), np.abs(dLs).astype(np.float64) @ np.abs(Wn).T)


def test_create_network_returns_the_reference_five_values():
    """steps.CtcHead.create_network = bilstm_ctc_net.py:31-52 after the recurrent layers: (logits, loss, model, prob,
    ler), each against the oracle chain (greedy decoder)."""
    from neuralasr_b200.steps import CtcHead
    from neuralasr_b200.utils import sparse_tuple_from
    from oracle import c_oracle
    rng = np.random.default_rng(9)
    B, T, K, C, Lmax = 4, 40, 500, 38, 8
    outputs = rng.standard_normal((B, T, K)).astype(np.float32)
    W = (rng.standard_normal((K, C)) * (3.0 / np.sqrt(K))).astype(np.float32)
    b = (rng.standard_normal(C) * 0.1).astype(np.float32)
    lens = rng.integers(1, Lmax + 1, size=B)
    labels = sparse_tuple_from(rng.integers(0, C - 1, size=(B, Lmax)), lens)
    seq = np.array([T, T, 31, T], np.int32)
    head = CtcHead(decoder="greedy")
    logits, loss, model, prob, ler = head.create_network(torch.from_numpy(outputs).cuda(), torch.from_numpy(W).cuda(),
                                                         torch.from_numpy(b).cuda(), labels, seq, B)
    torch.cuda.synchronize()
    want_logits = ao.tail(outputs, W, b, B)
    assert tuple(logits.shape) == (T, B, C) and np.abs(logits.cpu().numpy() - want_logits).max() < 1e-4
    x32 = np.ascontiguousarray(logits.cpu().numpy())        # decode on the kernel's own float32 logits: bit-exact ids
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    want_loss, _, _ = c_oracle.ctc_loss_grad(x32, labels[1], offs, seq, precision="f64")
    assert np.allclose(loss.per_utterance.cpu().numpy(), want_loss, rtol=1e-4)
    hv, ho, nsl = c_oracle.greedy_decode(x32, seq)
    assert np.array_equal(model.values.cpu().numpy(), hv)
    assert np.allclose(prob[:, 0].cpu().numpy(), nsl, rtol=1e-6)
    want_d, want_ler = c_oracle.edit_distance(hv, ho, labels[1], offs)
    assert np.array_equal(ler.distances.cpu().numpy(), want_d)
    assert abs(float(ler) - float(np.asarray(want_ler, np.float64).mean())) < 1e-6

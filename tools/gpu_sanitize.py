"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
narrow and wide throughput kernels, robust kernel, decode, label error rate, host-buffer call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import make_batch
from neuralasr_b200.networks import common
from neuralasr_b200 import host
from test_ctc_gpu import _triple

for kw in [dict(T=40, B=3, C=38, Lmax=12, mode="ragged"), dict(T=33, B=2, C=132, Lmax=20, mode="ragged"),
           dict(T=24, B=2, C=5, Lmax=3, mode="full", empty_row=False)]:
    g = make_batch(3, **kw)
    x = torch.from_numpy(g["logits"]).cuda()
    for path in (0, 1):
        common.debug_config(path, 0)
        loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"])
    common.debug_config(0, 0)
    dec, _ = common.decoding(x, g["seq_len"])
    ler = common.label_error_rate(dec, _triple(g))
    torch.cuda.synchronize()
    print(kw, float(loss.sum()), float(ler))
g = make_batch(4, T=40, B=3, C=38, Lmax=12, mode="ragged")
ctx = host.HostContext(0, 40, 3, 38, 12)
out = ctx.step(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"])
ctx.close()
print("host ok", out["loss"])

# the affine projection (all three kernels + the small-contraction kernel, aligned and misaligned rows) and the
# two-ended / one-ended edit distance
for rows, K, C in [(70, 500, 38), (33, 40, 38), (50, 700, 41), (17, 9, 3)]:
    H = torch.randn((rows, K), device="cuda")
    W = torch.randn((K, C), device="cuda")
    b = torch.randn((C,), device="cuda")
    y = common.affine_logits(H, W, b)
    dH, dW, db = common.affine_backward(H, W, torch.randn((rows, C), device="cuda"))
    big = torch.zeros((rows, K + 3), device="cuda")
    y2 = common.affine_logits(big[:, 1:K + 1], W, b)
    torch.cuda.synchronize()
    print("affine", rows, K, C, float(y.sum()), float(dW.sum()))
rng = np.random.default_rng(0)
for n, m in [(100, 40), (40, 300), (31, 5), (200, 600)]:
    hyp = torch.from_numpy(rng.integers(0, 30, size=(2, n))).cuda()
    hl = torch.tensor([n, n // 2], dtype=torch.int32, device="cuda")
    tv = rng.integers(0, 30, size=2 * m).astype(np.int32)
    idx = np.stack([np.repeat(np.arange(2), m), np.tile(np.arange(m), 2)], 1).astype(np.int64)
    d, ler = common.edit_distance(common.DecodedSparse(hyp, hl), (idx, tv, np.array([2, m], np.int64)))
    torch.cuda.synchronize()
    print("edit distance", n, m, d.tolist())

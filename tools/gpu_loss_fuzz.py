"""Randomised comparison of nasr_ctc_loss_grad with the C oracle over wide-vocabulary shapes: register-held rows,
streamed rows, odd lengths, batch-major and offset (misaligned) views.  Seeded."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import make_batch  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402
from oracle import c_oracle  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(777)
bad = handed = 0
for case in range(n_cases):
    C = int(rng.choice([65, 66, 100, 131, 256, 500, 1001, 1024, 1025, 1500, 2048, 3187, 4096, 6001, 8192]))
    Lmax = int(rng.integers(1, 200))
    T = int(rng.integers(max(2 * Lmax + 2, 17), 2 * Lmax + 120))
    B = int(rng.integers(1, 5))
    while T * B * C > 24_000_000:
        T = max(2 * Lmax + 2, T // 2)
        if T * B * C > 24_000_000:
            B = max(1, B - 1)
            Lmax = max(1, Lmax // 2)
    g = make_batch(5000 + case, T=T, B=B, C=C, Lmax=Lmax, mode=["ragged", "full", "tight"][case % 3],
                   peaky=bool(case % 2), empty_row=bool(case % 4 == 0))
    x = torch.from_numpy(g["logits"]).cuda()
    layout = case % 3
    if layout == 1:                       # batch-major storage, viewed time-major
        x = x.transpose(0, 1).contiguous().transpose(0, 1)
    elif layout == 2:                     # rows offset by one float inside a wider buffer: misaligned
        big = torch.zeros((T, B, C + 3), device="cuda")
        big[:, :, 1:C + 1] = x
        x = big[:, :, 1:C + 1]
    lab = (np.stack([np.repeat(np.arange(B), np.diff(g["label_offsets"])),
                     np.concatenate([np.arange(n) for n in np.diff(g["label_offsets"])]) if g["label_values"].size else
                     np.zeros(0, np.int64)], 1).astype(np.int64), g["label_values"],
           np.asarray([B, max(1, int(np.diff(g["label_offsets"]).max()))], np.int64))
    loss, grad, status = common.ctc_loss_and_grad(x, lab, g["seq_len"])
    handed += int((common.retry_flags(x.device, B) != 0).sum())
    wl, wg, ws = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    loss, grad, status = loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()
    fin = np.isfinite(wl)
    ok = np.array_equal(status, ws) and np.array_equal(np.isfinite(loss), fin)
    ok = ok and np.allclose(loss[fin], wl[fin], rtol=1e-4, atol=1e-5) and np.abs(grad - wg).max() <= 1e-4
    if not ok:
        bad += 1
        print("MISMATCH case %d: T=%d B=%d C=%d Lmax=%d layout=%d  max grad err %.2e" % (
            case, T, B, C, Lmax, layout, np.abs(grad - wg).max()), flush=True)
print("%d cases, %d mismatching, %d utterances handed to the retry kernel" % (n_cases, bad, handed))

"""Randomised comparison of nasr_ctc_loss_grad with the C oracle.  Seeded.
  python tools/gpu_loss_fuzz.py [n_cases]          wide vocabularies: register-held rows, streamed rows, odd lengths
  python tools/gpu_loss_fuzz.py [n_cases] narrow   C <= 64: every slots-per-lane build of the narrow kernel (transcripts
                                                   up to 638 labels), short inputs that go to the retry kernel
Both run batch-major and offset (misaligned) views beside the plain layout; with NASR_NARROW_F32=1 in the environment the
narrow mode exercises the float32 kernel of csrc/ctc_narrow.cu instead of the default."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import make_batch  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402
from oracle import c_oracle  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
narrow = len(sys.argv) > 2 and sys.argv[2] == "narrow"
rng = np.random.default_rng(4242 if narrow else 777)
bad = handed = 0
handed_by = {}   # why utterances went to the retry kernel: the shape class of their case
for case in range(n_cases):
    if narrow:
        C = int(rng.choice([2, 3, 5, 17, 29, 38, 40, 41, 63, 64]))
        # longest transcript of the batch picks the build: 2, 4, 5, 7, 10 or 20 slots per lane (63 .. 638 labels)
        Lmax = int(rng.choice([1, 9, 40, 63, 64, 127, 128, 159, 160, 200, 223, 224, 319, 320, 500, 638, 639, 700]))
        T = int(rng.integers(max(2 * Lmax + 2, 4), 2 * Lmax + 400))
        B = int(rng.integers(1, 9))
    else:
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
                   peaky=bool(case % 2), empty_row=bool(case % 4 == 0),
                   **({"repeat_p": 0.0 if C == 2 else 0.15} if narrow else {}))
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
    rf = common.retry_flags(x.device, B).cpu().numpy() != 0
    handed += int(rf.sum())
    if rf.any():
        lens = np.diff(g["label_offsets"])
        for b in np.nonzero(rf)[0]:
            why = ("transcript > 638 labels" if narrow and Lmax > 638 else
                   "fewer than 16 frames" if g["seq_len"][b] < 16 else
                   "empty transcript" if lens[b] == 0 else
                   "peaked rows (class ratio outside float range / range alarm)" if case % 2 else "regular")
            if why == "regular":
                slack = int(g["seq_len"][b]) - int(lens[b] + np.count_nonzero(np.diff(
                    g["label_values"][g["label_offsets"][b]:g["label_offsets"][b + 1]]) == 0))
                why = ("seq_len = frames needed (one alignment, p underflows the certificate)" if slack == 0 else
                       "2- or 3-class vocabulary (every second neighbour repeats: few alignments)" if C <= 3
                       else "regular (slack %d frames or more)" % (slack // 50 * 50))
                if len(sys.argv) > 3:
                    print("  case %d: T=%d B=%d C=%d Lmax=%d b=%d seq=%d L=%d" % (
                        case, T, B, C, Lmax, b, g["seq_len"][b], lens[b]))
            handed_by[why] = handed_by.get(why, 0) + 1
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
for why, n in sorted(handed_by.items(), key=lambda kv: -kv[1]):
    print("  handed over: %4d  %s" % (n, why))

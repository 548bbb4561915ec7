"""Streamed wide rows at a bandwidth-sized batch: python tools/gpu_stream_check.py [C] [B] [T] — prints ms, algorithmic
GB/s and the fraction of the measured HBM peak (for ncu: one call after two warm-ups)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import make_batch  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402

C, B, T = (int(a) for a in (sys.argv[1:4] + ["3000", "128", "800"][len(sys.argv) - 1:]))
g = make_batch(11, T, B, C, 150, mode="full", empty_row=False)
x = torch.from_numpy(g["logits"]).cuda()
lab = common.LabelsCSR(torch.from_numpy(g["label_values"]).cuda(), torch.from_numpy(g["label_offsets"]).cuda(), 150, B)
seq = torch.from_numpy(g["seq_len"]).cuda()
gbuf = torch.empty_like(x)
for _ in range(2):
    common.ctc_loss_and_grad(x, lab, seq, out_grad=gbuf)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
a.record()
for _ in range(n):
    loss, grad, status = common.ctc_loss_and_grad(x, lab, seq, out_grad=gbuf)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
alg = 12.0 * C * B * T
peak = 6543.4
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
handed = int((common.retry_flags(x.device, B) != 0).sum())
print("C=%d B=%d T=%d: %.3f ms, %.0f GB/s algorithmic = %.1f %% of %.0f GB/s; %d handed over; mean loss %.3f" % (
    C, B, T, ms, alg / ms / 1e6, 100 * alg / ms / 1e6 / peak, peak, handed, loss.mean().item()), flush=True)

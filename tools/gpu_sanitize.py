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

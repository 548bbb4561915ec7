"""kernel_ms of nasr_ctc_loss_grad at a bench workload for the library NASR_CTC_LIB points at (A/B runs on one box)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402

import os
if os.environ.get('NASR_DBG'):
    common.debug_config(int(os.environ['NASR_DBG'], 0), 0)
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
T, B, C = w["T"], w["B"], w["C"]
x, vals, offs, seq = bench.synth(w, 1234)
dev = torch.device("cuda", 0)
lens = np.diff(offs)
rows = np.repeat(np.arange(B), lens)
cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
lab = common.prepare_labels((np.stack([rows, cols], 1).astype(np.int64), vals, np.asarray([B, int(lens.max())])), dev)
seq_d = torch.from_numpy(seq).to(dev)
xs = [torch.from_numpy(x).to(dev) for _ in range(5)]
gs = [torch.empty_like(xs[0]) for _ in range(5)]
for i in range(5):
    common.ctc_loss_and_grad(xs[i], lab, seq_d, out_grad=gs[i])
torch.cuda.synchronize()
rf = common.retry_flags(dev, B).cpu().numpy()
print("utterances handed to the robust kernel: %d of %d, reasons %s" % (int((rf != 0).sum()), B, sorted(set(rf[rf != 0].tolist()))), flush=True)
res = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(50):
        common.ctc_loss_and_grad(xs[i % 5], lab, seq_d, out_grad=gs[i % 5])
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 50)
print("ms per call:", " ".join("%.4f" % r for r in res), flush=True)

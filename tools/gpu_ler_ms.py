"""ms per call of greedy decode + label error rate, and of the label error rate alone, at a bench workload."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from neuralasr_b200.networks import common
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
x, vals, offs, seq = bench.synth(w, 1234)
dev = torch.device("cuda", 0); B = w["B"]
lens = np.diff(offs); rows = np.repeat(np.arange(B), lens); cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
lab = common.prepare_labels((np.stack([rows, cols], 1).astype(np.int64), vals, np.asarray([B, int(lens.max())])), dev)
xs = torch.from_numpy(x).to(dev); seq_d = torch.from_numpy(seq).to(dev)
for _ in range(3):
    d, _ = common.decoding(xs, seq_d); common.edit_distance(d, lab)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    d, _ = common.decoding(xs, seq_d); common.edit_distance(d, lab)
b.record(); torch.cuda.synchronize()
print("decode + LER ms:", a.elapsed_time(b) / 20)
a.record()
for _ in range(20):
    common.edit_distance(d, lab)
b.record(); torch.cuda.synchronize()
print("LER alone ms:", a.elapsed_time(b) / 20, " mean hypothesis length %.0f" % float(d.hyp_len.float().mean()))

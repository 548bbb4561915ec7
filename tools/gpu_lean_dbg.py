"""Raw results of the throughput kernel alone (no retry) against the C oracle, per utterance and per frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
os.environ.setdefault('NASR_NARROW_F32', '1')
from conftest import make_batch
from neuralasr_b200.networks import common
from oracle import c_oracle

kw = dict(T=500, B=16, C=38, Lmax=100, mode="ragged")
seed = 4242
if len(sys.argv) > 1 and sys.argv[1] == "peaky":
    kw = dict(T=400, B=12, C=38, Lmax=60, mode="ragged", peaky=True)
if len(sys.argv) > 1 and sys.argv[1] == "peaky4":
    kw = dict(T=400, B=12, C=38, Lmax=100, mode="ragged", peaky=True)
if len(sys.argv) > 1 and sys.argv[1] == "rand2":
    kw = dict(T=400, B=12, C=38, Lmax=60, mode="ragged")
split = 0
if len(sys.argv) > 1 and sys.argv[1] == "uneven":
    kw = dict(T=120, B=6, C=38, Lmax=30, mode="ragged"); seed = 99; split = 104
if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
    kw = dict(T=1000, B=32, C=38, Lmax=200, mode="full"); seed = 7
g = make_batch(seed, **kw)
want_loss, want_grad, _ = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
dev = torch.device("cuda", 0)
B = kw["B"]
lens = np.diff(g["label_offsets"])
rows = np.repeat(np.arange(B), lens)
cols = np.arange(g["label_offsets"][-1]) - np.repeat(g["label_offsets"][:-1], lens)
lab = (np.stack([rows, cols], 1).astype(np.int64), g["label_values"], np.asarray([B, int(max(lens.max(), 1))]))
x = torch.from_numpy(g["logits"]).to(dev)
common.debug_config(2, split)
loss, grad, st = common.ctc_loss_and_grad(x, lab, g["seq_len"], out_grad=torch.zeros_like(x))
torch.cuda.synchronize()
fl = common.retry_flags(dev, B).cpu().numpy()
loss = loss.cpu().numpy(); grad = grad.cpu().numpy()
for b in range(B):
    Tb = int(g["seq_len"][b])
    err = np.abs(grad[:, b] - want_grad[:, b]).max(axis=1)
    badf = np.nonzero(~(err <= 1e-4))[0]
    print("b %2d Tb %4d L %3d flag %4d loss %.5f want %.5f  grad maxerr %.2e  bad frames %d %s" % (
        b, Tb, lens[b], fl[b], loss[b], want_loss[b], err.max(), len(badf),
        (str(badf[:6]) + ".." + str(badf[-6:])) if len(badf) else ""))
    if len(badf) and "-v" in sys.argv:
        t = badf[-1]
        print("   last bad frame", t, "got", grad[t, b][:8], "want", want_grad[t, b][:8])
        t = badf[0]
        print("   first bad frame", t, "got", grad[t, b][:8], "want", want_grad[t, b][:8])
        print("   bad frames:", badf.tolist()[:100])

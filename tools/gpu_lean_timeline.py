"""Per-chunk time stamps of the five warps of CTA 0 of the lean narrow kernel (NASR_TUNING=1 build)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from neuralasr_b200 import _lib
from neuralasr_b200.networks import common

w = dict(bench.WORKLOADS["cfg3"])
if len(sys.argv) > 1:
    w["B"] = int(sys.argv[1])
if len(sys.argv) > 2:
    w["Lmax"] = int(sys.argv[2])
T, B, C = w["T"], w["B"], w["C"]
x, vals, offs, seq = bench.synth(w, 1234)
dev = torch.device("cuda", 0)
lens = np.diff(offs)
rows = np.repeat(np.arange(B), lens)
cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
lab = common.prepare_labels((np.stack([rows, cols], 1).astype(np.int64), vals, np.asarray([B, int(lens.max())])), dev)
seq_d = torch.from_numpy(seq).to(dev)
xs = torch.from_numpy(x).to(dev)
g = torch.empty_like(xs)
for _ in range(3):
    common.ctc_loss_and_grad(xs, lab, seq_d, out_grad=g)
torch.cuda.synchronize()
prof = torch.zeros(3700 + 1024, dtype=torch.int64, device=dev)
_lib.load().nasr_debug_profile(ctypes.c_void_p(prof.data_ptr()))
common.ctc_loss_and_grad(xs, lab, seq_d, out_grad=g)
torch.cuda.synchronize()
_lib.load().nasr_debug_profile(ctypes.c_void_p(0))
p = prof.cpu().numpy()
ts = p[:3200].reshape(5, 160, 4)
t0 = ts[ts > 0].min()
names = ["R_F", "R_B", "PROD", "G_F", "G_B"]
for r in range(5):
    print("==", names[r])
    for c in range(160):
        if ts[r, c].max() == 0:
            continue
        a = ts[r, c] - t0
        extra = ""
        if r == 2:
            v = int(p[3200 + c]); extra = " q0=%d nq=%d" % (v >> 32, ((v & 0xffffffff) ^ 0x80000000) - 0x80000000)
        if c < 8 or 58 <= c < 72 or c > 118:
            print("%3d  t0 %8d  wait %6d  work %6d  tail %6d%s" % (c, a[0], a[1] - a[0], a[2] - a[1], a[3] - a[2], extra))

if "-w" in sys.argv:
    m = p[3400:3400 + 256].reshape(64, 4)
    by_sm = {}
    for bb in range(64):
        for wv in range(4):
            v = int(m[bb, wv]); sm = v >> 32; hw = (v >> 8) & 0xffff; role = v & 0xff
            by_sm.setdefault(sm, []).append((bb, wv, hw, names[role] if role < 3 else "GRAD"))
    for sm in sorted(by_sm)[:12]:
        print("SM", sm, by_sm[sm])

if "-e" in sys.argv:
    se = p[3700:3700 + 2 * B].reshape(B, 2)
    par = (se[:, 1] >> 62) & 1
    end = se[:, 1] & ((1 << 62) - 1)
    g0 = se[:, 0].min()
    st = (se[:, 0] - g0) / 1000.0; en = (end - g0) / 1000.0
    print("CTA start us: min %.1f max %.1f; end us: min %.1f median %.1f max %.1f" % (st.min(), st.max(), en.min(), np.median(en), en.max()))
    for q in (0, 1):
        m = par == q
        if m.any():
            print(" parity %d: n %d  duration us mean %.1f min %.1f max %.1f" % (q, m.sum(), (en - st)[m].mean(), (en - st)[m].min(), (en - st)[m].max()))
    order = np.argsort(en)
    print(" slowest CTAs:", [(int(i), int(par[i]), round(float(en[i] - st[i]), 1), int(lens[i])) for i in order[-8:]])
    print(" fastest CTAs:", [(int(i), int(par[i]), round(float(en[i] - st[i]), 1), int(lens[i])) for i in order[:8]])

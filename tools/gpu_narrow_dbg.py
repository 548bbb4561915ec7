"""Bring-up helper for the narrow kernel: one uneven-split case (for compute-sanitizer) or the indices of the cfg3
utterances that were handed to the robust kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import make_batch
from neuralasr_b200.networks import common
from test_ctc_gpu import _triple

dev = torch.device("cuda", 0)
which = sys.argv[1]
if which == "split":
    g = make_batch(hash("cfg1") % 1000, T=500, B=int(sys.argv[3]) if len(sys.argv) > 3 else 16, C=38, Lmax=100, mode="ragged")
    common.debug_config(2, int(sys.argv[2]))
    x = torch.from_numpy(g["logits"]).to(dev)
    loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"], out_grad=torch.zeros_like(x))
    torch.cuda.synchronize()
    print("ok", loss[:4].cpu().numpy(), common.retry_flags(dev, len(g["seq_len"])).cpu().numpy())
elif which == "one":
    import bench, ctypes
    from neuralasr_b200 import _lib
    from oracle import c_oracle
    w = bench.WORKLOADS["cfg3"]
    x, vals, offs, seq = bench.synth(w, 1234)
    b = int(sys.argv[2])
    lab1 = vals[offs[b]:offs[b + 1]]
    L = len(lab1)
    xs = np.ascontiguousarray(np.repeat(x[:, b:b + 1], 2, axis=1))
    idx = np.stack([np.repeat(np.arange(2, dtype=np.int64), L), np.tile(np.arange(L, dtype=np.int64), 2)], 1)
    lab = common.prepare_labels((idx, np.tile(lab1, 2), np.asarray([2, 200])), dev)
    dbg = torch.zeros(4096, dtype=torch.int32).pin_memory()
    _lib.load().nasr_debug_profile(ctypes.c_void_p(dbg.data_ptr()))
    common.debug_config(2, 0)
    xd = torch.from_numpy(xs).to(dev)
    loss, grad, st = common.ctc_loss_and_grad(xd, lab, torch.from_numpy(seq[:2].copy()).to(dev), out_grad=torch.zeros_like(xd))
    torch.cuda.synchronize()
    wl, wg, _ = c_oracle.ctc_loss_grad(xs, np.tile(lab1, 2), np.array([0, L, 2 * L], np.int32), seq[:2], precision="f64")
    print("retry", common.retry_flags(dev, 2).cpu().numpy(), "loss", loss.cpu().numpy(), "want", wl, "markers", dbg[:64].view(2, 8, 4)[0, :2].tolist())
    for nm, off, fl in (("A", 3000, 1), ("B", 3256, 1), ("E", 3512, 0), ("F", 3768, 1)):
        v = dbg[off:off + 256]
        v = v.view(torch.float32) if fl else v
        print("bwd", nm, "lanes 1,2:", v.view(8, 32)[:4, 1:3].T.numpy().ravel())
    term = dbg[64:64 + 512].view(torch.float32).view(16, 32).numpy(); kt = dbg[64 + 512:64 + 1024].view(16, 32).numpy(); lt = dbg[64 + 1024:64 + 1056].view(torch.float32).numpy()
    print("lane totals", lt)
    k, l = np.unravel_index(np.argmax(np.where(term > 0, kt + np.log2(np.maximum(term, 1e-45)), -1e9)), term.shape)
    print("max term at k", k, "lane", l, term[k, l], kt[k, l]); bad = np.argwhere(~np.isfinite(lt)); base = 64 + 1088
    for q in bad.ravel()[:1]:
        print("A", dbg[base:base + 256].view(torch.float32).view(8, 32)[:, q].numpy(), "pre", dbg[base + 256:base + 512].view(torch.float32).view(8, 32)[:, q].numpy(), "E", dbg[base + 512:base + 768].view(8, 32)[:, q].numpy(), "Eo", dbg[base + 768:base + 1024].view(8, 32)[:, q].numpy())
    print("bad lanes", bad.ravel(), [ (term[:, q], kt[:, q]) for q in bad.ravel()[:1]])
else:
    import bench
    w = bench.WORKLOADS["cfg3"]
    x, vals, offs, seq = bench.synth(w, 1234)
    B = w["B"]
    lens = np.diff(offs)
    rows = np.repeat(np.arange(B), lens)
    cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
    lab = common.prepare_labels((np.stack([rows, cols], 1).astype(np.int64), vals, np.asarray([B, int(lens.max())])), dev)
    common.debug_config(2, 0)
    xd = torch.from_numpy(x).to(dev)
    common.ctc_loss_and_grad(xd, lab, torch.from_numpy(seq).to(dev), out_grad=torch.zeros_like(xd))
    torch.cuda.synchronize()
    rf = common.retry_flags(dev, B).cpu().numpy()
    print("flagged:", [(int(b), int(rf[b]), int(lens[b])) for b in np.nonzero(rf)[0]])

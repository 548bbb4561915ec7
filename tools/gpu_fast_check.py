"""Bring-up check of the throughput CTC kernel on a GPU box: per-utterance errors against the C oracle
with the kernel running alone (debug path 2), the utterances it hands to the robust kernel, and timings.
The per-role cycle counts (`roles`, `roles5`, traces) and the ablation runs need a library built with the tuning hooks:
NASR_TUNING=1 python -m neuralasr_b200._build  (production builds compile them out)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import make_batch
from oracle import c_oracle
from neuralasr_b200.networks import common
from test_ctc_gpu import _triple

dev = torch.device("cuda", 0)

def run(name, path, split, **kw):
    g = make_batch(hash(name) % 1000, **kw)
    common.debug_config(path, split)
    x = torch.from_numpy(g["logits"]).to(dev)
    grad0 = torch.full_like(x, float("nan"))
    loss, grad, status = common.ctc_loss_and_grad(x, _triple(g), g["seq_len"], out_grad=grad0)
    torch.cuda.synchronize()
    B = kw["B"]
    retry = common.retry_flags(dev, B).cpu().numpy() if path != 1 else np.zeros(B, np.int32)
    wl, wg, ws = c_oracle.ctc_loss_grad(g["logits"], g["label_values"], g["label_offsets"], g["seq_len"], precision="f64")
    loss = loss.cpu().numpy(); grad = grad.cpu().numpy()
    ok = retry == 0 if path == 2 else np.ones(B, bool)
    fin = np.isfinite(wl) & ok
    lerr = np.abs(loss[fin] - wl[fin]) / np.maximum(np.abs(wl[fin]), 1e-3)
    gerr = np.array([np.nanmax(np.abs(grad[:, b] - wg[:, b])) if ok[b] else 0.0 for b in range(B)])
    nanb = [b for b in range(B) if ok[b] and not np.isfinite(grad[:, b]).all()]
    print("%-12s path=%d split=%-4d B=%-3d retried=%-3d reasons=%s  max loss rel err %.2e  max grad abs err %.2e  nan utts %s" % (
        name, path, split, B, int((retry != 0).sum()), sorted(set(retry[retry != 0].tolist())), lerr.max() if lerr.size else 0, gerr.max(), nanb[:8]))
    if gerr.max() > 1e-4 or (lerr.size and lerr.max() > 1e-4) or nanb:
        bb = int(np.argmax(gerr))
        print("   worst utt %d: Tb=%d L=%d loss %.6f want %.6f ; grad err by frame (first 12 bad frames): %s" % (
            bb, g["seq_len"][bb], np.diff(g["label_offsets"])[bb], loss[bb], wl[bb],
            np.nonzero(np.abs(grad[:, bb] - wg[:, bb]).max(-1) > 1e-4)[0][:12]))
    return g

cases = [
    ("small", dict(T=64, B=4, C=38, Lmax=10, mode="full", empty_row=False)),
    ("ragged", dict(T=120, B=9, C=38, Lmax=30, mode="ragged")),
    ("cfg1", dict(T=500, B=16, C=38, Lmax=100, mode="ragged")),
    ("cfg2", dict(T=800, B=64, C=38, Lmax=150, mode="ragged")),
    ("peaky", dict(T=400, B=12, C=38, Lmax=60, mode="ragged", peaky=True)),
    ("tight", dict(T=120, B=9, C=38, Lmax=50, mode="tight")),
    ("c64", dict(T=100, B=5, C=64, Lmax=20, mode="ragged")),
    ("c5", dict(T=40, B=3, C=5, Lmax=6, mode="ragged")),
    ("long", dict(T=1500, B=3, C=38, Lmax=300, mode="full")),
]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "parity"):
    for name, kw in cases:
        run(name, 2, 0, **kw)
    for split in (8, 16, 40, 56):
        run("small", 2, split, T=64, B=4, C=38, Lmax=10, mode="full", empty_row=False)
    run("cfg1", 2, 96, T=500, B=16, C=38, Lmax=100, mode="ragged")
    run("cfg1", 0, 0, T=500, B=16, C=38, Lmax=100, mode="ragged")
    common.debug_config(0, 0)
if which in ("all", "time"):
    g = make_batch(1234, T=1000, B=256, C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
    x = torch.from_numpy(g["logits"]).to(dev)
    lab = common.prepare_labels(_triple(g), dev)
    seq = torch.from_numpy(g["seq_len"]).to(dev)
    xs = [x.clone() for _ in range(4)]
    gs = [torch.empty_like(x) for _ in range(4)]
    for path in (2, 0, 1):
        common.debug_config(path, 0)
        for i in range(3):
            common.ctc_loss_and_grad(xs[i % 4], lab, seq, out_grad=gs[i % 4])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if path != 1 else 4
        a.record()
        for i in range(n):
            common.ctc_loss_and_grad(xs[i % 4], lab, seq, out_grad=gs[i % 4])
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print("cfg3 path %d: %.3f ms per step  -> %.2f%% of HBM peak (116.7 MB / 6543 GB/s = 17.8 us)" % (path, ms, 100 * 0.01784 / ms))
    common.debug_config(0, 0)
    run("cfg3", 2, 0, T=1000, B=256, C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
if which in ("all", "time", "roles", "roles5"):
    import ctypes
    from neuralasr_b200 import _lib
    lib = _lib.load()
    if which == "roles5":
        shape = dict(T=800, B=128, C=1024, Lmax=150, mode="full", Lmin=75, empty_row=False)
    else:
        shape = dict(T=1000, B=int(os.environ.get("NB", "256")), C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
    NB = shape["B"]
    g = make_batch(1234, **shape)
    x = torch.from_numpy(g["logits"]).to(dev)
    lab = common.prepare_labels(_triple(g), dev)
    seq = torch.from_numpy(g["seq_len"]).to(dev)
    prof = torch.zeros(NB * 16 * 4 + 4 * 200 * 8 * 2, dtype=torch.int64, device=dev)
    common.debug_config(2, 0)
    gr = torch.empty_like(x)
    common.ctc_loss_and_grad(x, lab, seq, out_grad=gr)
    lib.nasr_debug_profile(ctypes.c_void_p(prof.data_ptr()))
    common.ctc_loss_and_grad(x, lab, seq, out_grad=gr)
    torch.cuda.synchronize()
    lib.nasr_debug_profile(None)
    common.debug_config(0, 0)
    trace = prof.cpu().numpy()[NB * 64:].reshape(4, 200, 8, 2)
    np.save(os.path.join(ROOT, 'gpurun_out', 'r1_trace.npy'), trace)
    pr = prof.cpu().numpy()[:NB * 64].reshape(NB, 16, 4)
    np.save(os.path.join(ROOT, 'gpurun_out', 'r1_trace_roles.npy'), pr[:4, :8, 3])
    names = ["H_F", "H_B", "RC_F", "RC_B", "P_F", "P_B", "G_F", "G_B"]
    print("per-role cycles (mean over CTAs | max): work before meeting, work after meeting, total")
    live = pr[:, :, 2] > 0
    for r in range(8):
        sel = ((pr[:, :, 3] & 255) == r) & live
        w1, w2, tot = pr[:, :, 0][sel], pr[:, :, 1][sel], pr[:, :, 2][sel]
        print("  %-5s (%2d warps per CTA) phase1 %8.0f | %8d   phase2 %8.0f | %8d   total %8.0f | %8d" % (
            names[r], sel.sum() // NB, w1.mean(), w1.max(), w2.mean(), w2.max(), tot.mean(), tot.max()))
    smid = pr[:, 0, 3] >> 8
    import collections
    by = collections.defaultdict(list)
    for bb in range(NB):
        by[int(smid[bb])].append(bb)
    pairs = [v for v in by.values() if len(v) == 2]
    print("  SMs with two CTAs: %d" % len(pairs))
    tot = pr[:, 0, 2]
    print("  total cycles per CTA: mean %.0f max %d" % (tot.mean(), tot.max()))
if which == "ablate":
    g = make_batch(1234, T=1000, B=256, C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
    x = torch.from_numpy(g["logits"]).to(dev)
    lab = common.prepare_labels(_triple(g), dev)
    seq = torch.from_numpy(g["seq_len"]).to(dev)
    xs = [x.clone() for _ in range(4)]
    gs = [torch.empty_like(x) for _ in range(4)]
    for mask, name in [(0, "full"), (1, "no gradient warps"), (2, "no producers in phase 2"), (4, "no recompute"), (7, "recursion only"), (1 | 4, "no grad, no recompute")]:
        common.debug_config(2 | (mask << 8), 0)
        for i in range(3):
            common.ctc_loss_and_grad(xs[i % 4], lab, seq, out_grad=gs[i % 4])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            common.ctc_loss_and_grad(xs[i % 4], lab, seq, out_grad=gs[i % 4])
        b.record(); torch.cuda.synchronize()
        print("ablate %-28s %.3f ms" % (name, a.elapsed_time(b) / 20))
    # loss only = phase 1 only
    common.debug_config(2, 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        common.ctc_loss_and_grad(xs[i % 4], lab, seq, want_grad=False)
    torch.cuda.synchronize(); a.record()
    for i in range(20):
        common.ctc_loss_and_grad(xs[i % 4], lab, seq, want_grad=False)
    b.record(); torch.cuda.synchronize()
    print("loss only (phase 1 + meeting)        %.3f ms" % (a.elapsed_time(b) / 20))
    common.debug_config(0, 0)
if which == "bsweep":
    for NB in (64, 128, 148, 192, 256, 296):
        g = make_batch(1234, T=1000, B=NB, C=38, Lmax=200, mode="full", Lmin=100, empty_row=False)
        x = torch.from_numpy(g["logits"]).to(dev)
        lab = common.prepare_labels(_triple(g), dev)
        seq = torch.from_numpy(g["seq_len"]).to(dev)
        xs = [x.clone() for _ in range(6)]
        gs = [torch.empty_like(x) for _ in range(6)]
        for i in range(3):
            common.ctc_loss_and_grad(xs[i % 6], lab, seq, out_grad=gs[i % 6])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            common.ctc_loss_and_grad(xs[i % 6], lab, seq, out_grad=gs[i % 6])
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print("B=%3d: %.3f ms per call, %.2f us per utterance" % (NB, ms, 1e3 * ms / NB))

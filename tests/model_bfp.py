"""numpy model of the arithmetic of the fast CUDA kernel (csrc/ctc_fast.cu) — TEST INFRASTRUCTURE.

The kernel does not run TF's log-space recursion; it runs an algebraically
identical one chosen for the GPU.  This file restates that arithmetic lane by lane so that the
reformulation itself can be checked against the oracle on the CPU (tests/test_model_bfp.py) before any
GPU time is spent, and so that a GPU mismatch can be bisected.

Reformulation (all per utterance):
  * ratio units: every emission is divided by the frame's blank probability,
        R[t][c] = exp(x[t,c] - x[t,blank])  (blank -> 1),
    so blank states need no multiply; log p gets sum_t log y_blank(t) added back at the end;
  * pairs: slot i = j+1 holds (blank state 2j, label state 2j+1); slot 0 is a dummy, slot L+1 holds the
    final blank only.  Lane l owns slots [l*NL, (l+1)*NL);
  * the backward recursion is the same code on the reversed label string (mirrored slots
    i' = N-1-i), run over descending frames; "beta including the emission" is what it carries;
  * block floating point: lane values are fp64 with one int exponent per lane, renormalised every
    RESCALE frames (lane max -> [1,2));
  * meet in the middle: forward covers frames [0,M), backward covers [M,Tb) in phase 1; in phase 2 each
    continues through the other half against the other direction's rows, which are recomputed from
    checkpoints and kept only as the high 32 bits of each double;
  * posterior of label state = (pre-emission sum of the running direction) * (stored value of the other
    direction) / p; the blank's occupancy is 1 - sum of the label occupancies.
"""
from __future__ import annotations

import numpy as np

RESCALE = 8
GROWTH_BOUND = 500      # bits a lane max may grow between two rescales before the kernel gives up
Z_ALARM = 480           # Ea + Eb - ep above this: a flushed state (< 2^(Ea-1022)) times the largest
                        # possible partner (< 2^(Eb+1+GROWTH_BOUND)) over p could exceed 2^-34


def _hi_word(v):
    """Keep only the high 32 bits of each double (what the kernel stores for the other direction)."""
    a = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
    return (a & np.uint64(0xFFFFFFFF00000000)).view(np.float64)


class Direction:
    """One direction (forward or mirrored-backward) of the recursion in lane layout."""

    def __init__(self, lab, NL, nblk):
        L = len(lab)
        self.NL, self.nblk = NL, nblk
        N = NL * nblk
        assert N >= L + 2
        self.N, self.L = N, L
        self.col = np.full(N, -1, dtype=np.int64)         # class of the label in slot i (-1: dead)
        self.skip = np.zeros(N, dtype=np.float64)
        self.Ab = np.zeros(N)
        self.Al = np.zeros(N)
        self.E = np.zeros(nblk, dtype=np.int64)
        self.alarm = False

    @classmethod
    def forward(cls, lab, NL, nblk):
        d = cls(lab, NL, nblk)
        L = len(lab)
        d.col[1:L + 1] = lab
        if L > 1:
            d.skip[2:L + 1] = (np.asarray(lab[1:]) != np.asarray(lab[:-1])).astype(np.float64)
        d.first = 1
        d.Ab[1] = 1.0                                      # virtual row before frame 0
        return d

    @classmethod
    def backward(cls, lab, NL, nblk):
        d = cls(lab, NL, nblk)
        L = len(lab)
        pad = d.N - L - 1
        r = np.asarray(lab)[::-1]
        d.col[pad:pad + L] = r
        if L > 1:
            d.skip[pad + 1:pad + L] = (r[1:] != r[:-1]).astype(np.float64)
        d.first = pad
        d.Ab[pad] = 1.0                                    # virtual row after frame Tb-1
        return d

    def _lane(self, a):
        return a.reshape(self.nblk, self.NL)

    def rescale(self):
        Ab, Al = self._lane(self.Ab), self._lane(self.Al)
        m = np.maximum(Ab.max(1), Al.max(1))
        e = np.zeros(self.nblk, dtype=np.int64)
        nz = m > 0
        e[nz] = np.floor(np.log2(m[nz])).astype(np.int64)
        if np.any(e > GROWTH_BOUND):
            self.alarm = True
        with np.errstate(under="ignore"):
            Ab *= np.exp2(-e.astype(np.float64))[:, None]
            Al *= np.exp2(-e.astype(np.float64))[:, None]
        self.E += e
        # all-zero lanes adopt the exponent of the lane below (transitively)
        for l in range(1, self.nblk):
            if not nz[l]:
                self.E[l] = self.E[l - 1]

    def presums(self):
        """(nb, pre): next blank values and pre-emission label sums of the next frame."""
        NL = self.NL
        Ab, Al = self._lane(self.Ab), self._lane(self.Al)
        ll = np.zeros_like(Al)
        ll[:, 1:] = Al[:, :-1]
        d = np.zeros(self.nblk)
        d[1:] = (self.E[:-1] - self.E[1:]).astype(np.float64)
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            inflow = np.zeros(self.nblk)
            inflow[1:] = Al[:-1, NL - 1] * np.exp2(np.clip(d[1:], -1022, 1023)) * \
                np.exp2(np.clip(d[1:] - np.clip(d[1:], -1022, 1023), -1022, 1023))
        ll[:, 0] = inflow
        nb = Ab + ll
        pre = Al + Ab + self._lane(self.skip) * ll
        return nb.reshape(-1), pre.reshape(-1)

    def step(self, Rrow):
        """Advance one frame.  Rrow[c] = ratio emission of class c (blank column unused)."""
        nb, pre = self.presums()
        mult = np.where(self.col >= 0, Rrow[np.maximum(self.col, 0)], 0.0)
        self.Ab = nb
        self.Al = mult * pre
        return pre

    def snapshot(self):
        return self.Ab.copy(), self.Al.copy(), self.E.copy()

    def restore(self, snap):
        self.Ab, self.Al, self.E = snap[0].copy(), snap[1].copy(), snap[2].copy()


def ctc_loss_grad_model(x, lab, blank, NL=8, K=16, split=None, f32_exp=True):
    """One utterance, x [Tb, C] float32 logits.  Returns (loss, grad[Tb,C], info)."""
    x = np.asarray(x, dtype=np.float32)
    Tb, C = x.shape
    lab = np.asarray(lab, dtype=np.int64)
    L = len(lab)
    nblk = (L + 2 + NL - 1) // NL
    N = NL * nblk
    # producer: softmax pieces and ratio table, float32 arithmetic like the kernel
    m = x.max(1, keepdims=True)
    n = np.exp((x - m).astype(np.float32)).astype(np.float32)
    s = n.sum(1, dtype=np.float32)
    inv_s = (np.float32(1) / s).astype(np.float32)
    with np.errstate(divide="ignore", over="ignore"):
        R = (n * (np.float32(1) / n[:, blank:blank + 1])).astype(np.float32).astype(np.float64)
    logyb = (x[:, blank] - m[:, 0]).astype(np.float32).astype(np.float64) - np.log(s.astype(np.float64))
    # split point, a multiple of K
    if split is None:
        M = Tb
    else:
        M = int(split)
    M = max(0, min(Tb, M))
    fw = Direction.forward(lab, NL, nblk)
    bw = Direction.backward(lab, NL, nblk)
    ck_f, ck_b = {}, {}
    # ---- phase 1 ----
    for t in range(0, M):
        if t % RESCALE == 0:
            fw.rescale()
        if t % K == 0:
            ck_f[t] = fw.snapshot()
        fw.step(R[t])
    for t in range(Tb - 1, M - 1, -1):
        if (t + 1) % RESCALE == 0:
            bw.rescale()
        if (t + 1) % K == 0 or t == Tb - 1:
            ck_b[t] = bw.snapshot()
        bw.step(R[t])
    # ---- p at the meeting point: alpha row of frame M-1 against the backward presums of frame M-1 ----
    fw_meet, bw_meet = fw.snapshot(), bw.snapshot()
    nb_b, pre_b = bw.presums()
    # label in forward slot i <-> backward slot N-1-i; forward blank of slot i (state 2(i-1)) <-> backward
    # blank slot N-i (state 2(L-j'') with j'' = slot - pad)
    idx = np.arange(N)
    lane = idx // NL
    Eb_for_label = bw.E[(N - 1 - idx) // NL]
    with np.errstate(over="ignore", under="ignore"):
        termL = fw.Al * pre_b[N - 1 - idx]
        eL = fw.E[lane] + Eb_for_label
        bidx = N - idx
        okb = (bidx >= 0) & (bidx < N)
        termB = np.where(okb, fw.Ab * nb_b[np.clip(bidx, 0, N - 1)], 0.0)
        eB = fw.E[lane] + bw.E[np.clip(bidx, 0, N - 1) // NL]
    terms = np.concatenate([termL, termB])
    exps = np.concatenate([eL, eB]).astype(np.float64)
    nzm = terms > 0
    info = {"alarm": False, "M": M}
    if not np.any(nzm):
        return np.inf, None, {"alarm": True, "M": M}
    lg = np.log2(terms[nzm]) + exps[nzm]
    top = lg.max()
    ptot = np.exp2(lg - top).sum()
    log2p = top + np.log2(ptot)
    ep = int(np.floor(log2p))
    mp = float(np.exp2(log2p - ep))
    loss = -(log2p * np.log(2.0) + logyb.sum())
    # ---- phase 2 ----
    occ = np.zeros((Tb, C))                 # label occupancies by class
    alarm = fw.alarm or bw.alarm

    def run_phase2(cont, other_ckpts, frames_segments, other_is_fw):
        nonlocal alarm
        for seg in frames_segments:          # list of frames in the order `cont` walks them
            # recompute the other direction's rows for this segment from its checkpoint
            oth = Direction.forward(lab, NL, nblk) if other_is_fw else Direction.backward(lab, NL, nblk)
            order = sorted(seg) if other_is_fw else sorted(seg, reverse=True)
            oth.restore(other_ckpts[order[0]])
            rows, rowE = {}, {}
            for t in order:
                bound = t if other_is_fw else t + 1
                if bound % RESCALE == 0 and t != order[0]:
                    oth.rescale()
                oth.step(R[t])
                rows[t] = _hi_word(oth.Al)
                rowE[t] = oth.E.copy()
            for t in seg:
                if (t % RESCALE == 0) if not other_is_fw else ((t + 1) % RESCALE == 0):
                    cont.rescale()
                _, pre = cont.presums()
                st = rows[t][N - 1 - idx]
                Est = rowE[t][(N - 1 - idx) // NL]
                k = cont.E[lane] + Est - ep
                kk = k[(cont.col >= 0) & (pre > 0) & (st > 0)]
                if kk.size:
                    info["kmax"] = max(info.get("kmax", -10**9), int(kk.max()))
                if np.any((k > Z_ALARM) & (cont.col >= 0) & (pre > 0) & (st > 0)):
                    alarm = True
                with np.errstate(over="ignore", under="ignore"):
                    g = pre * st * np.exp2(np.clip(k, -2000, 2000).astype(np.float64)) / mp
                g = g.astype(np.float32).astype(np.float64)
                ok = cont.col >= 0
                np.add.at(occ[t], cont.col[ok], g[ok])
                cont.step(R[t])
            alarm = alarm or oth.alarm

    def segments(lo, hi, ascending):
        segs = []
        t = lo
        while t < hi:
            e = min(hi, (t // K + 1) * K)
            segs.append(list(range(t, e)))
            t = e
        if not ascending:
            segs = [s[::-1] for s in segs[::-1]]
        return segs

    # forward continues over [M, Tb) against backward rows; backward continues over [0, M) against forward rows
    fw.restore(fw_meet)
    run_phase2(fw, ck_b, segments(M, Tb, True), other_is_fw=False)
    bw.restore(bw_meet)
    run_phase2(bw, ck_f, segments(0, M, False), other_is_fw=True)
    alarm = alarm or fw.alarm or bw.alarm
    y = n.astype(np.float64) * inv_s.astype(np.float64)[:, None]
    occ[:, blank] = 1.0 - occ.sum(1)
    grad = y - occ
    info["alarm"] = alarm
    return float(loss), grad, info

"""numpy model of the arithmetic of the narrow-vocabulary CUDA kernel (csrc/ctc_narrow.cu) — TEST INFRASTRUCTURE.

The kernel does not run TF's log-space recursion (reference call site
networks/tfnetwork.py:58-59); it runs an algebraically identical one chosen for the GPU.  This file restates that
arithmetic slot by slot in float32 so that the reformulation can be checked against the oracle on the CPU
(tests/test_model_f32.py) before any GPU time is spent, and so that a GPU mismatch can be bisected.

Reformulation (all per utterance):
  * ratio units: R[t][c] = exp(x[t,c] - x[t,blank]); blank states need no multiply; log p gets
    sum_t log y_blank(t) added back at the end;
  * slots: slot i = j+1 holds (blank state 2j, label state 2j+1); slot 0 is a dummy, slot L+1 holds the final
    blank only; the backward recursion is the same code on the reversed label string (mirrored slots N-1-i) over
    descending frames and carries "beta including the emission";
  * float32 values with ONE INTEGER EXPONENT PER SLOT: true value = stored * 2^E[i].  The value that crosses
    from slot i-1 into slot i is multiplied by F[i] = 2^(E[i-1]-E[i]) inside the fused multiply-add that consumes it;
  * phase 1 (forward over [0,M), backward over [M,Tb)): every SEG frames each slot is renormalised so that its larger
    state has biased exponent TB, subject to E[i] >= E[i-1] - GCAP over slots that hold mass (so F <= 2^GCAP);
    empty slots adopt the exponent of the slot below;
  * p from the two directions at the meeting point; the state is divided by the mantissa of p;
  * phase 2 (each direction continues through the other half): the slot exponent is SET to
    e_p - E_other[N-1-i] ("posterior gauge"), so that stored(own, pre-emission) * stored(other) IS the posterior of
    the label state; the other direction's values are recomputed from the checkpoints of phase 1;
  * a-posteriori certificate: flushes only ever remove mass, so the totals alpha(T-1)[final states]/p and
    beta(0)[first states]/p are <= 1 up to rounding and fall short of 1 by exactly the posterior mass the other
    direction lost in phase 1 (and this one in phase 2).  Both must be within TOL of 1, else the utterance is
    handed to the robust kernel.
"""
from __future__ import annotations

import numpy as np

SEG = 8           # frames between renormalisations / checkpoints
TB = 60           # biased exponent the larger state of a slot is brought to (2^-67)
BIG = np.float32(2.0 ** 90)  # label values saturate here: with F <= 2^GCAP nothing can reach inf, so no NaN can arise
GCAP = 30         # a slot with mass sits at most this far below the slot with mass beneath it
PSHIFT = 64       # the posterior buffer holds posterior * 2^-PSHIFT (two values near 2^(TB-127) are multiplied)
FCLAMP = GCAP     # largest exponent of a transfer factor
ENEG = -(1 << 28)
TOL = 3e-5
f32 = np.float32


def _pow2(e):
    """2^e as float32 with the flush / overflow behaviour of building the float from its exponent field."""
    e = np.asarray(e, dtype=np.int64)
    out = np.zeros(e.shape, dtype=np.float32)
    ok = (e >= -126) & (e <= 127)
    out[ok] = np.exp2(e[ok].astype(np.float64)).astype(np.float32)
    out[e > 127] = np.inf
    return out


class Dir:
    def __init__(self, lab, N, backward):
        lab = np.asarray(lab, dtype=np.int64)
        L = len(lab)
        assert N >= L + 2
        self.N, self.L = N, L
        self.NL = N // 32
        self.col = np.full(N, -1, dtype=np.int64)
        self.skip = np.zeros(N, dtype=np.float32)
        self.Ab = np.zeros(N, dtype=np.float32)
        self.Al = np.zeros(N, dtype=np.float32)
        self.E = np.full(N, ENEG, dtype=np.int64)
        self.F = np.zeros(N, dtype=np.float32)
        self.alarm = False
        self.maxexp = -999
        if not backward:
            self.col[1:L + 1] = lab
            if L > 1:
                self.skip[2:L + 1] = (lab[1:] != lab[:-1]).astype(np.float32)
            first = 1
        else:
            pad = N - L - 1
            r = lab[::-1]
            self.col[pad:pad + L] = r
            if L > 1:
                self.skip[pad + 1:pad + L] = (r[1:] != r[:-1]).astype(np.float32)
            first = pad
        self.first = first
        self.Ab[first] = f32(2.0) ** f32(TB - 127)   # virtual row before the first frame, already in slot scale
        self.E[first] = -(TB - 127)
        self.E[first + 1:] = self.E[first]
        self._set_F()

    def _set_F(self):
        d = np.zeros(self.N, dtype=np.int64)
        d[1:] = self.E[:-1] - self.E[1:]
        dead = np.zeros(self.N, dtype=bool)
        dead[0] = True
        dead[1:] = (self.E[:-1] == ENEG) | (self.E[1:] == ENEG)
        self.F = np.where(dead, f32(0), _pow2(np.minimum(np.where(dead, 0, d), FCLAMP))).astype(np.float32)

    def snapshot(self):
        return self.Ab.copy(), self.Al.copy(), self.E.copy()

    def restore(self, s):
        self.Ab, self.Al, self.E = s[0].copy(), s[1].copy(), s[2].copy()
        self._set_F()

    def presums(self):
        alp = np.zeros(self.N, dtype=np.float32)
        alp[1:] = self.Al[:-1]
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            nb = (self.Ab + self.F * alp).astype(np.float32)
            q = ((self.Al + self.Ab) + (self.skip * self.F) * alp).astype(np.float32)
        return nb, q

    def step(self, Rrow):
        nb, q = self.presums()
        mult = np.where(self.col >= 0, Rrow[np.maximum(self.col, 0)], f32(0)).astype(np.float32)
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            self.Al = np.minimum((q * mult).astype(np.float32), BIG)
        self.Ab = nb
        return q

    def _scale(self, Enew):
        d = np.where((Enew == ENEG) | (self.E == ENEG), 0, Enew - self.E)   # stored *= 2^-d
        # two multiplications: one power of two cannot span [2^-252, 2^252]
        dd = np.clip(-d, -252, 252)
        h = dd // 2
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            s1, s2 = _pow2(h), _pow2(dd - h)
            self.Ab = ((self.Ab * s1).astype(np.float32) * s2).astype(np.float32)
            self.Al = ((self.Al * s1).astype(np.float32) * s2).astype(np.float32)
        self.E = Enew.copy()
        self._set_F()

    def rescale_phase1(self):
        m = np.maximum(self.Ab, self.Al)
        bits = m.view(np.uint32).astype(np.int64)
        ex = bits >> 23
        if np.any(ex >= 255) or np.any(np.isnan(self.Ab)) or np.any(np.isnan(self.Al)):
            self.alarm = True
        self.maxexp = max(self.maxexp, int(ex.max()))
        # a slot whose values have decayed into the denormals still holds mass: exponent from m * 2^64
        with np.errstate(over="ignore", under="ignore"):
            exd = ((m * f32(2.0 ** 64)).astype(np.float32).view(np.uint32).astype(np.int64) >> 23) - 64
        ex = np.where(ex == 0, exd, ex)
        nz = (m > 0) & (self.E != ENEG)
        own = np.where(nz, self.E + ex - TB, ENEG)
        Enew = np.full(self.N, ENEG, dtype=np.int64)
        NL = self.NL
        nl = self.N // NL
        # pass 0: every lane runs its own chain from the neighbour lane's last exponent AS OF THE PREVIOUS RESCALE;
        # passes 1, 2: the chain is raised against the neighbour's value of the previous pass (one shuffle each, no scan
        # over lanes: a raise that has to travel further than two lanes in one rescale is left to the clamp of F)
        for ps in range(3):
            src = self.E if ps == 0 else Enew.copy()
            for l in range(nl):
                prev = src[l * NL - 1] if l else ENEG
                for i in range(l * NL, l * NL + NL):
                    if nz[i]:
                        cand = own[i] if ps == 0 else Enew[i]
                        prev = cand if prev == ENEG else max(cand, prev - GCAP)
                    Enew[i] = prev
        self._scale(Enew)

    def rescale_gauge(self, Eo, ep):
        """Posterior gauge: E[i] = ep - Eo[N-1-i]; where the label's partner slot is dead use the blank's partner
        (slot N-i); where both are dead the slot is dead."""
        N = self.N
        idx = np.arange(N)
        pl = Eo[N - 1 - idx]
        pb = np.where(idx >= 1, Eo[np.clip(N - idx, 0, N - 1)], ENEG)
        part = np.where(pl != ENEG, pl, pb)
        Enew = np.where(part != ENEG, ep - part, ENEG)
        # a dead slot keeps nothing
        deadnow = Enew == ENEG
        self.Ab[deadnow] = 0
        self.Al[deadnow] = 0
        ex = (np.maximum(self.Ab, self.Al).view(np.uint32).astype(np.int64)) >> 23
        if np.any(ex >= 255) or np.any(np.isnan(self.Ab)) or np.any(np.isnan(self.Al)):
            self.alarm = True
        self.maxexp = max(self.maxexp, int(ex.max()))
        self._scale(Enew)


def ctc_loss_grad_model(x, lab, blank, NL=8, split=None):
    """One utterance, x [Tb, C] float32 logits.  Returns (loss, grad[Tb,C], info)."""
    x = np.asarray(x, dtype=np.float32)
    Tb, C = x.shape
    lab = np.asarray(lab, dtype=np.int64)
    L = len(lab)
    N = NL * 32
    assert L + 2 <= N and Tb >= 2 * SEG
    with np.errstate(over="ignore", under="ignore"):
        R = np.exp2(((x - x[:, blank:blank + 1]) * f32(1.4426950408889634)).astype(np.float32)).astype(np.float32)
    rs = R.sum(1, dtype=np.float32)
    yb = (f32(1) / rs).astype(np.float32)
    logyb = -np.log(rs.astype(np.float64))
    info = {"alarm": False}
    if not (np.all(np.isfinite(rs)) and R.min() >= f32(1.1754944e-38)):
        return None, None, {"alarm": True, "why": "emission"}
    M = SEG * ((Tb + SEG) // (2 * SEG)) if split is None else (int(split) // SEG) * SEG
    M = max(SEG, min(M, SEG * ((Tb - 1) // SEG)))
    info["M"] = M
    fw, bw = Dir(lab, N, False), Dir(lab, N, True)
    ck_f, ck_b = {}, {}
    # ---- phase 1 ----
    for t in range(M):
        if t % SEG == 0:
            if t:
                fw.rescale_phase1()
            ck_f[t // SEG] = fw.snapshot()
        fw.step(R[t])
    nB = Tb - M
    for tau in range(nB):
        if tau % SEG == 0:
            if tau:
                bw.rescale_phase1()
            ck_b[tau // SEG] = bw.snapshot()
        bw.step(R[Tb - 1 - tau])
    # ---- meeting: p = sum_u alpha(u, M-1) * beta_TF(u, M-1) ----
    fw.rescale_phase1()
    bw.rescale_phase1()
    nb_b, pre_b = bw.presums()
    idx = np.arange(N)
    okb = idx >= 1
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        tl = fw.Al.astype(np.float64) * pre_b[N - 1 - idx].astype(np.float64)
        el = fw.E + bw.E[N - 1 - idx]
        tb_ = np.where(okb, fw.Ab.astype(np.float64) * nb_b[np.clip(N - idx, 0, N - 1)].astype(np.float64), 0.0)
        eb = fw.E + bw.E[np.clip(N - idx, 0, N - 1)]
    terms = np.concatenate([tl, tb_])
    exps = np.concatenate([el, eb])
    live = (terms > 0) & (exps > ENEG // 2)
    if not np.any(live):
        return None, None, {"alarm": True, "why": "p"}
    lg = np.log2(terms[live]) + exps[live]
    top = int(np.floor(lg.max()))
    tot = np.exp2(lg - top).astype(np.float32).sum(dtype=np.float32)      # float32 accumulation like the kernel
    log2p = top + np.log2(float(tot))
    ep = int(np.floor(log2p))
    mp = f32(np.exp2(log2p - ep))
    loss = -(log2p * np.log(2.0) + logyb.sum())
    occ = np.zeros((Tb, C), dtype=np.float64)
    inv_mp = f32(1) / mp
    tot_max = 0.0

    def phase2(cont, other_ckpts, n_other, other_backward):
        """cont walks the other direction's phase-1 frames from the meeting point outwards."""
        nonlocal tot_max
        cont.Ab = (cont.Ab * inv_mp).astype(np.float32)
        cont.Al = (cont.Al * inv_mp).astype(np.float32)
        nseg = (n_other + SEG - 1) // SEG
        for s in range(nseg - 1, -1, -1):
            tau0, tau1 = s * SEG, min(n_other, s * SEG + SEG)
            oth = Dir(lab, N, other_backward)
            oth.restore(other_ckpts[s])
            rows = []
            for tau in range(tau0, tau1):
                t = Tb - 1 - tau if other_backward else tau
                oth.step(R[t])
                rows.append((t, oth.Al.copy()))
            cont.rescale_phase1()
            Eo = other_ckpts[s][2]
            k = np.where((cont.E == ENEG) | (Eo[N - 1 - idx] == ENEG), -200, cont.E + Eo[N - 1 - idx] - ep)
            kk = np.clip(k - PSHIFT, -252, 252)                # posteriors are kept as post * 2^-PSHIFT
            c1, c2 = _pow2(kk // 2), _pow2(kk - kk // 2)      # split so that neither partial product leaves the normal range
            c1 = np.where(k - PSHIFT < -252, f32(0), c1)       # beyond the reach of two factors: the posterior is zero
            for t, orow in rows[::-1]:
                _, q = cont.presums()
                with np.errstate(over="ignore", under="ignore", invalid="ignore"):
                    post = ((q * c1).astype(np.float32) * (orow[N - 1 - idx] * c2).astype(np.float32)).astype(np.float32)
                ok = cont.col >= 0
                np.add.at(occ[t], cont.col[ok], post[ok].astype(np.float64) * 2.0 ** PSHIFT)
                tot_max = max(tot_max, float(post[ok].sum(dtype=np.float32)) * 2.0 ** PSHIFT)
                cont.step(R[t])
        return cont

    fw = phase2(fw, ck_b, nB, True)
    bw = phase2(bw, ck_f, M, False)
    # ---- certificate ----
    def end_total(d, slot_blank, slot_label):
        v = 0.0
        if d.E[slot_blank] != ENEG:
            v += float(d.Ab[slot_blank]) * 2.0 ** float(np.clip(d.E[slot_blank] - ep, -300, 300))
        if L > 0 and d.E[slot_label] != ENEG:
            v += float(d.Al[slot_label]) * 2.0 ** float(np.clip(d.E[slot_label] - ep, -300, 300))
        return v
    tf_end = end_total(fw, L + 1, L)
    tb_end = end_total(bw, N - 1, N - 2)
    info.update(tot_fwd=tf_end, tot_bwd=tb_end, tot_max=tot_max, maxexp=max(fw.maxexp, bw.maxexp))
    bad = not (abs(tf_end - 1) < TOL and abs(tb_end - 1) < TOL and tot_max < 1 + 1e-4)
    info["alarm"] = bad or fw.alarm or bw.alarm
    y = R.astype(np.float64) * yb.astype(np.float64)[:, None]
    occ[:, blank] = 1.0 - occ.sum(1)
    return float(loss), y - occ, info

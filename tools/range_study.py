"""Dynamic-range study for the block-floating-point CTC recursion (design aid, not shipped code).

For random utterances of a workload it computes exact log2 alpha / beta lattices in fp64 log space and
reports, per group of NL consecutive (blank,label) pairs and per frame,
    k = max_group log2 alpha(.,t) + max_group log2 beta(.,t) - log2 p
i.e. how far above p the product of the group maxima is.  A flushed fp32 value is < 2^-126 of its group
maximum, so the damage of a flush is bounded by 2^(k-126)."""
import sys
import numpy as np

def lattices(x, lab, blank):
    T, C = x.shape
    L = len(lab)
    U = 2 * L + 1
    lp = x - x.max(1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(1, keepdims=True))
    ext = np.full(U, blank); ext[1::2] = lab
    skip = np.zeros(U, bool); skip[3::2] = lab[1:] != lab[:-1]
    em = lp[:, ext]                       # [T,U]
    NEG = -np.inf
    a = np.full((T, U), NEG); a[0, 0] = em[0, 0]
    if U > 1: a[0, 1] = em[0, 1]
    for t in range(1, T):
        p = a[t - 1]
        s1 = np.concatenate([[NEG], p[:-1]])
        s2 = np.concatenate([[NEG, NEG], p[:-2]]); s2 = np.where(skip, s2, NEG)
        a[t] = em[t] + np.logaddexp(np.logaddexp(p, s1), s2)
    # beta hat: includes emission at t
    b = np.full((T, U), NEG); b[T - 1, U - 1] = em[T - 1, U - 1]
    if U > 1: b[T - 1, U - 2] = em[T - 1, U - 2]
    skipn = np.zeros(U, bool); skipn[:-2] = skip[2:]
    for t in range(T - 2, -1, -1):
        p = b[t + 1]
        s1 = np.concatenate([p[1:], [NEG]])
        s2 = np.concatenate([p[2:], [NEG, NEG]]); s2 = np.where(skipn, s2, NEG)
        b[t] = em[t] + np.logaddexp(np.logaddexp(p, s1), s2)
    logp = np.logaddexp(a[T - 1, U - 1], a[T - 1, U - 2] if U > 1 else NEG)
    return a / np.log(2), b / np.log(2), em / np.log(2), logp / np.log(2)

def study(T, L, C, NL, seed, scale=3.0, peaky=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, C)) * scale
    lab = rng.integers(0, C - 1, size=L)
    for i in range(1, L):
        if rng.random() < 0.1: lab[i] = lab[i - 1]
    if peaky:
        x = rng.standard_normal((T, C))
        x[:, C - 1] += 8
        pos = np.sort(rng.choice(T, size=L, replace=False))
        for i, t in enumerate(pos):
            x[t, lab[i]] += 16
    a, b, em, logp = lattices(x, lab, C - 1)
    U = 2 * L + 1
    post = a + b - em - logp            # log2 posterior
    # groups of NL pairs = 2*NL states, slot 0 is a dummy -> states start at offset: pair j = states (2j,2j+1) in slot j+1
    npairs = L + 1
    nslots = npairs + 1
    ngrp = (nslots + NL - 1) // NL
    kmax = -1e9
    worst = None
    with np.errstate(invalid="ignore"):
        for g in range(ngrp):
            s0 = max(0, 2 * (g * NL - 1)); s1 = min(U, 2 * ((g + 1) * NL - 1))
            if s1 <= s0: continue
            ga = a[:, s0:s1].max(1); gb = (b - em)[:, s0:s1].max(1)
            k = ga + gb - logp
            k = k[np.isfinite(k)]
            if k.size and k.max() > kmax:
                kmax = k.max(); worst = g
    # also: how far below group max are states that carry posterior > 2^-24 ?
    need = 0.0
    for g in range(ngrp):
        s0 = max(0, 2 * (g * NL - 1)); s1 = min(U, 2 * ((g + 1) * NL - 1))
        if s1 <= s0: continue
        ga = a[:, s0:s1].max(1, keepdims=True)
        rel = a[:, s0:s1] - ga
        m = post[:, s0:s1] > -24
        if m.any(): need = min(need, rel[m].min())
    return kmax, worst, need, -logp * np.log(2)

if __name__ == "__main__":
    T, L, C, NL = [int(v) for v in sys.argv[1:5]]
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    peaky = len(sys.argv) > 6 and sys.argv[6] == "peaky"
    for seed in range(n):
        k, g, need, loss = study(T, L, C, NL, seed, peaky=peaky)
        print("seed %d: kmax=%.1f (group %d)  most-negative rel. exponent of a state with posterior>2^-24: %.1f  loss=%.1f" % (seed, k, g, need, loss))

"""HOST-buffer step with ordinary (pageable) numpy arrays against the context's pinned staging (cfg3)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from neuralasr_b200 import host  # noqa: E402

w = bench.WORKLOADS["cfg3"]
T, B, C = w["T"], w["B"], w["C"]
x, vals, offs, seq = bench.synth(w, 1234)
gl = np.full(B, 1.0 / B, np.float32)
ctx = host.HostContext(0, T, B, C, w["Lmax"])
pin = ctx.pinned_logits[: x.size].reshape(T, B, C)
pin[...] = x
gout = np.empty_like(x)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


print("pinned in, pinned out : %.2f ms" % timed(lambda: ctx.step(pin, vals, offs, seq, grad_loss=gl, want_decode=False)))
print("pageable in, pinned out: %.2f ms" % timed(lambda: ctx.step(x, vals, offs, seq, grad_loss=gl, want_decode=False)))
print("pageable in and out    : %.2f ms" % timed(lambda: ctx.step(x, vals, offs, seq, grad_loss=gl, want_decode=False,
                                                                    grad_out=gout)))
ctx.close()

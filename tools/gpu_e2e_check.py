"""e2e (HOST buffers through nasr_host_ctc_step) at cfg3 for the current NASR_HOST_BLOCKS / NASR_HOST_STREAMS."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from neuralasr_b200 import host  # noqa: E402
from oracle import c_oracle  # noqa: E402

w = bench.WORKLOADS["cfg3"]
T, B, C = w["T"], w["B"], w["C"]
x, vals, offs, seq = bench.synth(w, 1234)
ctx = host.HostContext(0, T, B, C, w["Lmax"])
pin = ctx.pinned_logits[: x.size].reshape(T, B, C)
pin[...] = x
gl = np.full(B, 1.0 / B, np.float32)
for _ in range(3):
    out = ctx.step(pin, vals, offs, seq, grad_loss=gl, want_decode=False)
t0 = time.perf_counter()
n = 30
for _ in range(n):
    out = ctx.step(pin, vals, offs, seq, grad_loss=gl, want_decode=False)
dt = (time.perf_counter() - t0) / n
print("blocks=%s streams=%s: %.3f ms per step, %.3e frames/s" % (
    os.environ.get("NASR_HOST_BLOCKS", "default"), os.environ.get("NASR_HOST_STREAMS", "default"), dt * 1e3,
    T * B / dt), flush=True)
if len(sys.argv) > 1:
    keys = out.keys() if isinstance(out, dict) else None
    loss = out["loss"] if keys else out[0]
    grad = out["grad"] if keys else out[1]
    want_loss, want_grad, _ = c_oracle.ctc_loss_grad(x[:, :8], vals[: offs[8]], offs[:9], seq[:8], precision="f64")
    print("check: max rel loss err %.2e, max abs grad err %.2e" % (
        np.abs(loss[:8] - want_loss).max() / np.abs(want_loss).max(),
        np.abs(np.asarray(grad).reshape(T, B, C)[:, :8] * B - want_grad).max()), flush=True)
ctx.close()

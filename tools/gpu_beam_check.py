"""Time the beam search kernel on the GPU box (random and peaky logits) and print per-call milliseconds."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from neuralasr_b200.networks import common  # noqa: E402


def peaky(T, B, C, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((T, B, C), device="cuda", generator=g)
    # planted alignment: a class held for 2 frames at a time, blank-dominated
    cls = torch.randint(0, C - 1, (T // 2 + 1, B), device="cuda", generator=g).repeat_interleave(2, 0)[:T]
    isb = torch.rand((T // 2 + 1, B), device="cuda", generator=g).repeat_interleave(2, 0)[:T] < 0.5
    cls = torch.where(isb, torch.full_like(cls, C - 1), cls)
    x.scatter_add_(2, cls.unsqueeze(-1), torch.full((T, B, 1), 8.0, device="cuda"))
    return x


def bench(name, x, W=100, reps=3):
    T, B, C = x.shape
    seq = np.full(B, T, np.int32)
    common.beam_decoding(x, seq, beam_width=W)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        dec, lp = common.beam_decoding(x, seq, beam_width=W)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print("%-28s T=%d B=%d C=%d W=%d  %.3f ms  %.3e frames/s  mean len %.1f  mean logp %.2f" % (
        name, T, B, C, W, ms, T * B / ms * 1e3, dec[0].hyp_len.float().mean().item(), lp.mean().item()), flush=True)


if __name__ == "__main__":
    for (T, B, C) in [(800, 64, 38), (1000, 256, 38)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        bench("random*3", torch.randn((T, B, C), device="cuda", generator=g) * 3)
        bench("peaky", peaky(T, B, C))
    bench("peaky wide", peaky(800, 128, 1024))
    g = torch.Generator(device="cuda").manual_seed(1)
    bench("random*3 wide", torch.randn((200, 128, 1024), device="cuda", generator=g) * 3)


def phases(name, x, W=100):
    """Cycles per frame of each phase of the frame loop (CTA 0), through nasr_debug_profile.  Needs a library built
    with the tuning hooks (NASR_TUNING=1 python -m neuralasr_b200._build); a production build reports zeros."""
    import ctypes
    from neuralasr_b200 import _lib
    lib = _lib.load()
    T, B, C = x.shape
    buf = torch.zeros(B * 16 * 4 + 4 * 200 * 8 * 2, dtype=torch.int64, device="cuda")
    lib.nasr_debug_profile(ctypes.c_void_p(buf.data_ptr()))
    common.beam_decoding(x, np.full(B, T, np.int32), beam_width=W)
    torch.cuda.synchronize()
    lib.nasr_debug_profile(None)
    ph = buf[:8].cpu().numpy() / T
    names = ["update", "lists", "keys", "select", "admit", "build", "parents"]
    print("%-20s cycles/frame: " % name + "  ".join("%s %.0f" % (k, v) for k, v in zip(names, ph)) +
          "  total %.0f" % ph.sum(), flush=True)


if __name__ == "__main__":
    g = torch.Generator(device="cuda").manual_seed(1)
    phases("random*3 B=64", torch.randn((800, 64, 38), device="cuda", generator=g) * 3)
    phases("peaky B=64", peaky(800, 64, 38))
    phases("peaky B=256", peaky(1000, 256, 38))
    phases("peaky wide", peaky(200, 128, 1024))

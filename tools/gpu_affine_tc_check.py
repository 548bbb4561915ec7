"""The tcgen05 forward of the affine projection (NASR_AFFINE_TC=1) against the float64 oracle, then its time."""
import os
import sys
os.environ["NASR_AFFINE_TC"] = "1"
import numpy as np
import torch
sys.path.insert(0, ".")
from neuralasr_b200.networks import common
from oracle import affine_oracle as ao

worst = 0.0
for rows, K, C in [(64, 32, 8), (128, 64, 38), (200, 256, 40), (1000, 500, 38), (4099, 500, 38), (70, 700, 5)]:
    rng = np.random.default_rng(rows + K)
    H = rng.standard_normal((rows, K)).astype(np.float32)
    W = (rng.standard_normal((K, C)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32)
    got = common.affine_logits(torch.from_numpy(H).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    want = ao.affine_logits(H, W, b)
    bound = np.abs(H).astype(np.float64) @ np.abs(W) + np.abs(b)
    rel = float((np.abs(got.cpu().numpy() - want) / bound).max())
    worst = max(worst, rel)
    print("rows %d K %d C %d: max |err| / (|H|.|W|) = %.3e  max |err| = %.3e" % (
        rows, K, C, rel, float(np.abs(got.cpu().numpy() - want).max())), flush=True)
print("worst", worst)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    rows, K, C = 256000, 500, 38
    g = torch.Generator(device="cuda").manual_seed(0)
    Hs = [torch.randn((rows, K), device="cuda", generator=g) for _ in range(2)]
    W = torch.randn((K, C), device="cuda", generator=g) / K ** 0.5
    b = torch.zeros((C,), device="cuda")
    out = torch.empty((rows, C), device="cuda")
    for i in range(3):
        common.affine_logits(Hs[i & 1], W, b, out=out)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        common.affine_logits(Hs[i & 1], W, b, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    print("tcgen05 forward %.4f ms  %.0f GB/s algorithmic" % (ms, 4 * (rows * K + rows * C) / ms / 1e6))

"""ms per call and fraction of the HBM peak of the affine projection kernels at a bench workload's row count
(rows = B*T, K = 500), beside torch's own float32 matmul (TF32 off) on the same tensors."""
import json
import sys
import torch
sys.path.insert(0, ".")
import bench
from neuralasr_b200.networks import common
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
import os
rows, K, C = w["B"] * w["T"], int(os.environ.get("AFFINE_K", "500")), w["C"]
peak = bench.hbm_peak()[0] if hasattr(bench, "hbm_peak") else 6543.4
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
Hs = [torch.randn((rows, K), device="cuda", generator=g) for _ in range(2)]   # 2 x 512 MB > L2
W = torch.randn((K, C), device="cuda", generator=g) / K ** 0.5
b = torch.zeros((C,), device="cuda")
dL = torch.randn((rows, C), device="cuda", generator=g)
out = torch.empty((rows, C), device="cuda")


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n


res = {"rows": rows, "K": K, "C": C, "hbm_peak_gbs": peak}
fwd_bytes = 4 * (rows * K + rows * C + K * C)
ms = timed(lambda i: common.affine_logits(Hs[i & 1], W, b, out=out))
res["forward"] = {"ms": ms, "gbs": fwd_bytes / ms / 1e6, "frac": fwd_bytes / ms / 1e6 / peak}
ms = timed(lambda i: torch.addmm(b, Hs[i & 1], W, out=out))
res["torch_forward_fp32"] = {"ms": ms, "gbs": fwd_bytes / ms / 1e6}
ms = timed(lambda i: common.affine_backward(Hs[i & 1], W, dL, True, False, False))
by = 4 * (rows * K + rows * C)
res["backward_dH"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
ms = timed(lambda i: common.affine_backward(Hs[i & 1], W, dL, False, True, True))
res["backward_dW_db"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
ms = timed(lambda i: (torch.mm(dL, W.t()), torch.mm(Hs[i & 1].t(), dL), dL.sum(0)))
res["torch_backward_fp32"] = {"ms": ms}
print(json.dumps(res))

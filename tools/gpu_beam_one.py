"""One beam search call on a peaky batch (for ncu): python tools/gpu_beam_one.py [B] [T] [C]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from gpu_beam_check import peaky  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402

B, T, C = (int(a) for a in (sys.argv[1:4] + ["148", "200", "38"][len(sys.argv) - 1:]))
x = peaky(T, B, C)
for _ in range(2):
    dec, lp = common.beam_decoding(x, np.full(B, T, np.int32))
torch.cuda.synchronize()
print("ok", dec[0].hyp_len.float().mean().item(), lp.mean().item())

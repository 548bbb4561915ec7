"""Floor of the HOST-buffer path on this box: what N ranks can move when each does nothing but
copy one batch of logits host->device and one batch of gradients device->host per step, from / to pinned memory.
Run alone or under torchrun.  Three patterns per rank:
  contiguous : one 4*T*B*C-byte cudaMemcpyAsync each way
  blocks     : eight column slabs [T, B/8, C] each way as pitched 2-D copies (what nasr_host_ctc_step issues)
  blocks-1d  : the same slabs staged contiguously, one linear copy each
H2D and D2H run on two streams at once.  Prints ms per step (max over ranks) and the aggregate GB/s per direction."""
import os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T, B, C = 1000, 256, 38
nbytes = 4 * T * B * C
h_in = torch.empty((T, B, C), dtype=torch.float32).pin_memory()
h_out = torch.empty((T, B, C), dtype=torch.float32).pin_memory()
d_in = torch.empty((T, B, C), dtype=torch.float32, device=dev)
d_out = torch.zeros((T, B, C), dtype=torch.float32, device=dev)
h_in.normal_()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
nb = 8
# contiguous staging of the slabs: [nb][T, B/nb, C]
h_in_b = torch.empty((nb, T, B // nb, C), dtype=torch.float32).pin_memory()
h_out_b = torch.empty((nb, T, B // nb, C), dtype=torch.float32).pin_memory()
d_in_b = torch.empty((nb, T, B // nb, C), dtype=torch.float32, device=dev)
d_out_b = torch.zeros((nb, T, B // nb, C), dtype=torch.float32, device=dev)


def step(pattern):
    with torch.cuda.stream(s_in):
        if pattern == "contiguous":
            d_in.copy_(h_in, non_blocking=True)
        elif pattern == "blocks":
            for k in range(nb):
                sl = slice(k * B // nb, (k + 1) * B // nb)
                d_in[:, sl, :].copy_(h_in[:, sl, :], non_blocking=True)
        else:
            for k in range(nb):
                d_in_b[k].copy_(h_in_b[k], non_blocking=True)
    with torch.cuda.stream(s_out):
        if pattern == "contiguous":
            h_out.copy_(d_out, non_blocking=True)
        elif pattern == "blocks":
            for k in range(nb):
                sl = slice(k * B // nb, (k + 1) * B // nb)
                h_out[:, sl, :].copy_(d_out[:, sl, :], non_blocking=True)
        else:
            for k in range(nb):
                h_out_b[k].copy_(d_out_b[k], non_blocking=True)


for pattern in ("contiguous", "blocks", "blocks-1d"):
    for _ in range(3):
        step(pattern)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        step(pattern)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) * 1e3
        print("%d rank(s) %-10s: %.3f ms per step (max over ranks) = %.1f GB/s per direction aggregate, %.3e frames/s"
              % (world, pattern, ms, world * nbytes / (ms * 1e-3) / 1e9, world * T * B / (ms * 1e-3)), flush=True)
if world > 1:
    dist.destroy_process_group()

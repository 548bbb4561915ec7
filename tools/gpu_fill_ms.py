import torch
x=[torch.empty((256000,500),device='cuda') for _ in range(2)]
for i in range(3): x[i&1].fill_(1.0)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20): x[i&1].fill_(float(i))
b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b)/20
print("fill 512 MB: %.4f ms = %.0f GB/s"%(ms, 512e6/ms/1e6))

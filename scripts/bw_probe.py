import torch, time
n = 1 << 30
x = torch.empty(n, dtype=torch.int32, device="cuda")   # 4 GB
y = torch.empty(n, dtype=torch.int32, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.zero_()); print(f"write-only (memset 4GB): {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: x.fill_(3)); print(f"write-only (fill 4GB): {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy 4GB->4GB: {2*4.295/ms*1e3:.0f} GB/s (r+w)")
xf = x.view(torch.float32)
ms = t(lambda: torch.sum(xf)); print(f"read-only (sum 4GB): {4.295/ms*1e3:.0f} GB/s")

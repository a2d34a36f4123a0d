import torch, time
x = torch.empty(8 * 1024**3 // 4, dtype=torch.int32, device="cuda").random_(0, 100)
for fn_name, fn in (("sum", lambda: x.sum()), ("max", lambda: x.max()), ("copy", lambda: x[: x.numel() // 2].copy_(x[x.numel() // 2:]))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = x.numel() * 4
    print(fn_name, f"{ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")

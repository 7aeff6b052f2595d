import torch, time
for mb in (1, 7, 50):
    a = torch.empty(mb * 1000_000, dtype=torch.uint8).pin_memory()
    b = torch.empty(mb * 1000_000, dtype=torch.uint8, device="cuda")
    for _ in range(3): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): b.copy_(a, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(mb, "MB pinned H2D:", mb * 10 / e0.elapsed_time(e1), "GB/s")

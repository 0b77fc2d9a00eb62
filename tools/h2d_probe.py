import torch, time
x = torch.empty(36_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for chunks in (1, 4, 16, 64):
    n = x.numel() // chunks
    for _ in range(3):
        for c in range(chunks): d[c*n:(c+1)*n].copy_(x[c*n:(c+1)*n], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for c in range(chunks): d[c*n:(c+1)*n].copy_(x[c*n:(c+1)*n], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"chunks {chunks}: {ms:.3f} ms  {36e6/ms/1e6:.1f} GB/s")

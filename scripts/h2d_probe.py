import torch, time, os
print("cpus", os.cpu_count(), torch.get_num_threads())
x = torch.rand((1024, 36, 2048)).pin_memory()
y = torch.empty((1024, 36, 2048), dtype=torch.bfloat16).pin_memory()
for nt in (4, 8, 16, 32):
    torch.set_num_threads(nt)
    y.copy_(x)
    t = time.perf_counter()
    for _ in range(5): y.copy_(x)
    dt = (time.perf_counter() - t) / 5
    print(f"cpu f32->bf16 threads={nt}: {dt*1e3:.2f} ms  ({x.numel()*6/dt/1e9:.1f} GB/s)")
d32 = torch.empty_like(x, device="cuda"); d16 = torch.empty((1024, 36, 2048), dtype=torch.bfloat16, device="cuda")
for name, src, dst in (("f32", x, d32), ("bf16", y, d16)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"H2D pinned {name}: {ms:.2f} ms  {src.numel()*src.element_size()/ms/1e6:.1f} GB/s")

import os, time, torch
print("conf", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_NO_CUDA_MEMORY_CACHING"))
def stat(): 
    s = torch.cuda.memory_stats(); return s.get("num_device_alloc"), s.get("num_device_free"), s["reserved_bytes.all.current"] >> 20
x = torch.empty(1, device="cuda"); torch.cuda.synchronize()
for label, hold in (("drop", False), ("hold one", True)):
    r = None
    t0 = time.perf_counter()
    for i in range(20):
        t = torch.empty((16, 4096, 4096), device="cuda")
        if hold: r = t
        del t
    torch.cuda.synchronize()
    print(label, f"{(time.perf_counter()-t0)/20*1e6:.1f} us/call", stat())
for n in (1 << 20, 1 << 26, 1 << 28, 1 << 30):
    t0 = time.perf_counter()
    for i in range(20):
        t = torch.empty(n, dtype=torch.uint8, device="cuda"); del t
    print(n >> 20, "MiB", f"{(time.perf_counter()-t0)/20*1e6:.1f} us/call", stat())

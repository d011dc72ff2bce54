import torch, time
# effective L2 capacity for cyclic re-reads: bandwidth of summing a buffer of X MB repeatedly
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size / 2**20, "MB")
for mb in (8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 128, 256):
    x = torch.empty(mb * 2**20 // 4, dtype=torch.float32, device="cuda").normal_()
    for _ in range(3): x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): x.sum()
    e1.record(); torch.cuda.synchronize()
    print(mb, "MB:", round(mb * 2**20 * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e12, 2), "TB/s")
